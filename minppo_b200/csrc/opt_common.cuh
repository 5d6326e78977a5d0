// minppo_b200 -- device helpers shared by the optimizer kernels (adam.cu, dwopt.cuh).
#pragma once

#include "common.cuh"
#include "minppo_internal.h"

namespace minppo {

constexpr int OPT_THREADS = 512;                 // 512 x <= 42 regs: co-resident with a fused-step CTA under PDL
constexpr int OPT_EPT = 8;                    // max elements per thread (registers)

template <int NTHREADS = 512>
MINPPO_DEVINL float block_sum(float v, float* scratch /*[32]*/) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x < 32) {
    s = threadIdx.x < (NTHREADS / 32) ? scratch[threadIdx.x] : 0.f;
    s = warp_sum(s);
  }
  __syncthreads();
  return s;                                   // valid in warp 0
}

// Sense-free grid barrier on a monotonically increasing 64-bit counter.  All blocks of the
// grid are co-resident (grid <= #SMs, one block per SM).  A bounded spin turns a scheduling
// surprise into an error flag instead of a hung GPU.
MINPPO_DEVINL void grid_barrier(unsigned long long* counter, int* err_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long old = atomicAdd(counter, 1ULL);
    const unsigned long long target = (old / gridDim.x + 1ULL) * gridDim.x;
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned long long*>(counter) < target) {
      if (clock64() - t0 > 4000000000LL) { atomicExch(err_flag, MINPPO_ERR_BARRIER); break; }
    }
    __threadfence();
  }
  __syncthreads();
}

// fixed-order sum of `nparts` partials, 8 loads in flight
MINPPO_DEVINL float sum_partials(const float* __restrict__ src, int nparts, size_t stride) {
  float acc = 0.f;
  int p = 0;
  for (; p + 8 <= nparts; p += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + static_cast<size_t>(p + u) * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  for (; p < nparts; ++p) acc += __ldcg(src + static_cast<size_t>(p) * stride);
  return acc;
}

MINPPO_DEVINL int find_leaf_idx(const OptArgs& a, int i) {
  int l = 0;
  while (l + 1 < a.nleaves && i >= a.leaf[l + 1].offset) ++l;
  return l;
}


// One Adam step on one element (optax.scale_by_adam + scale(-lr), train.py:116-124), after the
// global-norm clip (optax.clip_by_global_norm: g unchanged if gnorm < max_norm, else g / gnorm * max_norm).
struct AdamScalars { float gnorm, lr, c1, c2; bool trigger; };
MINPPO_DEVINL void adam_element(const OptArgs& a, const AdamScalars& sc, float g, float& p, float& m, float& v) {
  if (!sc.trigger) g = (g / sc.gnorm) * a.max_norm;
  m = a.one_minus_b1 * g + a.b1 * m;
  v = a.one_minus_b2 * (g * g) + a.b2 * v;
  const float u = (m / sc.c1) / (sqrtf(v / sc.c2 + a.eps_root) + a.eps);
  p = p + (-sc.lr) * u;
}
// learning rate of this step (train.py:98-101 annealed, else opt.lr) and the Adam bias corrections
MINPPO_DEVINL void step_scalars(const OptArgs& a, int count, float& lr, float& c1, float& c2) {
  if (a.anneal) {
    const float frac = 1.0f - static_cast<float>(count / a.anneal_div) / static_cast<float>(a.num_updates);
    lr = a.lr * frac;
  } else {
    lr = a.lr;
  }
  const float cnt1 = static_cast<float>(count + 1);
  c1 = 1.0f - powf(a.b1, cnt1);
  c2 = 1.0f - powf(a.b2, cnt1);
}
// bf16 images of a hidden kernel element (what the tcgen05 GEMMs read)
MINPPO_DEVINL void write_images(const OptLeaf& L, int i, float p) {
  if (L.img_t || L.img_n) {
    const int e = i - L.offset;
    const int r = e / L.cols, c = e % L.cols;              // kernel [in=r][out=c]
    const __nv_bfloat16 b = __float2bfloat16_rn(p);
    if (L.img_t) L.img_t[static_cast<size_t>(c) * L.ld_t + r] = b;
    if (L.img_n) L.img_n[static_cast<size_t>(r) * L.ld_n + c] = b;
  }
  if (L.img_w2) {                                          // head kernel [in=r][out=j] -> kernel^T hi / lo, SW128
    const int e = i - L.offset;
    const int r = e / L.cols, j = e % L.cols;
    uint32_t hi, lo;
    split_bf16(p, hi, lo);
    const uint32_t off = sw16_off(j, r);
    *reinterpret_cast<uint16_t*>(L.img_w2 + off) = static_cast<uint16_t>(hi);
    *reinterpret_cast<uint16_t*>(L.img_w2 + 8192 + off) = static_cast<uint16_t>(lo);
  }
}

}  // namespace minppo
