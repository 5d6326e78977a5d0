// minppo_b200 -- Threefry-2x32/20 and the two JAX bit-stream layouts (jax/_src/prng.py), shared by the
// permutation generator (prng_sort.cu; /root/reference/minppo/train.py:252,258) and the policy sampler
// (policy.cu; train.py:158-159).  Bit-exact with oracle/threefry.py.
#pragma once

#include <stdint.h>

#include "../../include/minppo_b200.h"

namespace minppo {

// ---------------------------------------------------------------------------------------
// Threefry-2x32, 20 rounds
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& o0,
                                             uint32_t& o1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int g = 0; g < 5; ++g) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl32(x1, R[g & 1][i]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + static_cast<uint32_t>(g + 1);
  }
  o0 = x0;
  o1 = x1;
}

// split(key) -> (first, second) child keys
__host__ __device__ inline void key_split(const uint32_t key[2], int mode, uint32_t a[2], uint32_t b[2]) {
  if (mode == MINPPO_PRNG_LEGACY) {
    // threefry_2x32(key, iota(4)): pairs (0,2) and (1,3); output = concat(o0[0:2], o1[0:2]) -> [[o0_0,o0_1],[o1_0,o1_1]]
    uint32_t p0, q0, p1, q1;
    threefry2x32(key[0], key[1], 0u, 2u, p0, q0);
    threefry2x32(key[0], key[1], 1u, 3u, p1, q1);
    a[0] = p0; a[1] = p1; b[0] = q0; b[1] = q1;
  } else {
    threefry2x32(key[0], key[1], 0u, 0u, a[0], a[1]);
    threefry2x32(key[0], key[1], 0u, 1u, b[0], b[1]);
  }
}

// random_bits(key, 32, (n,))[i]
__device__ __forceinline__ uint32_t random_bits_at(const uint32_t key[2], int mode, uint32_t i, uint32_t n) {
  uint32_t o0, o1;
  if (mode == MINPPO_PRNG_LEGACY) {
    const uint32_t h = (n + 1u) >> 1;                 // padded half length
    if (i < h) {
      const uint32_t hi = i + h;
      threefry2x32(key[0], key[1], i, hi < n ? hi : 0u, o0, o1);   // pad counter is 0
      return o0;
    }
    threefry2x32(key[0], key[1], i - h, i, o0, o1);
    return o1;
  }
  threefry2x32(key[0], key[1], 0u, i, o0, o1);
  return o0 ^ o1;
}

}  // namespace minppo
