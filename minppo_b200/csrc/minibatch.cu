// minppo_b200 -- minibatch bookkeeping between the permutation and the learner steps.
//
// The reference materialises a shuffled copy of every trajectory leaf per epoch
// (/root/reference/minppo/train.py:260-265).  Here nothing is copied: minibatch (e, k) is the
// index list perm[e][k*mb : (k+1)*mb] and rows are gathered on the fly by the GEMM producer.
//
// compact_rows: with env-sharded ranks every rank computes the GLOBAL
//   permutation and keeps, per minibatch, the entries whose env falls in its shard, in
//   permutation order, as LOCAL flat indices t*Nl + (n - n0).  world_size == 1 keeps all rows.
// adv_stats: the reference normalises advantages per minibatch inside the loss
//   (train.py:235: (gae - gae.mean()) / (gae.std() + 1e-8)).  The permutations depend only on
//   the key chain, so (sum, centred second moment) of all E*M minibatches are computed up front.
#include "common.cuh"
#include "minppo_internal.h"

namespace minppo {

constexpr int MB_THREADS = 256;

// one block per minibatch; ordered stream compaction in strips of MB_THREADS
__global__ void __launch_bounds__(MB_THREADS) compact_rows_kernel(const int32_t* __restrict__ perms,
                                                                  int32_t* __restrict__ rowidx,
                                                                  int32_t* __restrict__ counts, int M, long long B,
                                                                  int mb, int cap, int N, int n0, int Nl) {
  __shared__ int warp_tot[MB_THREADS / 32];
  __shared__ int base_s;
  const int s = blockIdx.x;                       // e * M + k
  const int e = s / M, k = s % M;
  const int32_t* src = perms + static_cast<size_t>(e) * B + static_cast<size_t>(k) * mb;
  int32_t* dst = rowidx + static_cast<size_t>(s) * cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < mb; i0 += MB_THREADS) {
    const int i = i0 + threadIdx.x;
    int local = -1;
    if (i < mb) {
      const int flat = src[i];
      const int t = flat / N, n = flat - t * N;
      if (n >= n0 && n < n0 + Nl) local = t * Nl + (n - n0);
    }
    const unsigned m = __ballot_sync(0xffffffffu, local >= 0);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    const int pos = off + __popc(m & ((1u << lane) - 1u));
    if (local >= 0 && pos < cap) dst[pos] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < MB_THREADS / 32; ++w) tot += warp_tot[w];
      base_s += tot;
    }
    __syncthreads();
  }
  const int count = base_s;
  if (threadIdx.x == 0) counts[s] = count;          // > cap is reported by the host-visible check
  for (int j = min(count, cap) + threadIdx.x; j < cap; j += MB_THREADS) dst[j] = 0;
}

// pass 0: stats[s] = sum adv ; pass 1: stats[EM+s] = sum (adv - stats[s]/mb)^2   (owned rows)
__global__ void __launch_bounds__(MB_THREADS) adv_stats_kernel(const float* __restrict__ adv,
                                                               const int32_t* __restrict__ rowidx,
                                                               const int32_t* __restrict__ counts,
                                                               float* __restrict__ stats, int EM, int cap,
                                                               int mb, int pass) {
  __shared__ float wsum[MB_THREADS / 32];
  const int s = blockIdx.x;
  const int count = min(counts[s], cap);
  const int32_t* idx = rowidx + static_cast<size_t>(s) * cap;
  const float mean = pass ? stats[s] / static_cast<float>(mb) : 0.f;
  float acc = 0.f;
  for (int j = threadIdx.x; j < count; j += MB_THREADS) {
    const float d = adv[idx[j]] - mean;
    acc += pass ? d * d : d;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < MB_THREADS / 32; ++w) t += wsum[w];
    stats[pass * EM + s] = t;
  }
}

int compact_rows_launch(const int32_t* perms, int32_t* rowidx, int32_t* counts, int E, int M, long long B, int mb,
                        int cap, int N, int n0, int Nl, cudaStream_t stream) {
  compact_rows_kernel<<<E * M, MB_THREADS, 0, stream>>>(perms, rowidx, counts, M, B, mb, cap, N, n0, Nl);
  return cudaGetLastError() == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
}

int adv_stats_launch(const float* adv, const int32_t* rowidx, const int32_t* counts, float* stats, int EM, int cap,
                     int mb, int pass, cudaStream_t stream) {
  adv_stats_kernel<<<EM, MB_THREADS, 0, stream>>>(adv, rowidx, counts, stats, EM, cap, mb, pass);
  return cudaGetLastError() == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
}

}  // namespace minppo
