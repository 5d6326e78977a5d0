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

// one block per minibatch; ordered stream compaction in strips of CR_THREADS x CR_ITEMS consecutive entries (one
// __syncthreads per strip: the warp totals are double-buffered).  The GLOBAL minibatch grows with the number of ranks
// (weak scaling: 65,536 entries per minibatch at 8 GPUs), so this scan is sized for that, not for one rank's rows.
constexpr int CR_THREADS = 1024;
constexpr int CR_ITEMS = 4;
__global__ void __launch_bounds__(CR_THREADS) compact_rows_kernel(const int32_t* __restrict__ perms,
                                                                  int32_t* __restrict__ rowidx,
                                                                  int32_t* __restrict__ counts, int M, long long B,
                                                                  int mb, int cap, int N, int n0, int Nl,
                                                                  int* __restrict__ err_flag) {
  __shared__ int warp_tot[2][CR_THREADS / 32];
  griddep_wait();                                 // launched with programmatic serialization (launch_chain)
  griddep_launch();
  const int s = blockIdx.x;                       // e * M + k
  const int e = s / M, k = s % M;
  const int32_t* src = perms + static_cast<size_t>(e) * B + static_cast<size_t>(k) * mb;
  int32_t* dst = rowidx + static_cast<size_t>(s) * cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;                                   // rows kept so far (identical in every thread)
  for (int i0 = 0, it = 0; i0 < mb; i0 += CR_THREADS * CR_ITEMS, ++it) {
    int loc[CR_ITEMS];
    int c = 0;
#pragma unroll
    for (int j = 0; j < CR_ITEMS; ++j) {
      const int i = i0 + static_cast<int>(threadIdx.x) * CR_ITEMS + j;
      loc[j] = -1;
      if (i < mb) {
        const int flat = src[i];
        const int t = flat / N, n = flat - t * N;
        if (n >= n0 && n < n0 + Nl) { loc[j] = t * Nl + (n - n0); ++c; }
      }
    }
    int inc = c;                                  // inclusive scan of the per-thread counts over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[it & 1][warp] = inc;
    __syncthreads();
    const int wv = warp_tot[it & 1][lane];        // CR_THREADS / 32 == 32 warp totals, one per lane
    const int before = __reduce_add_sync(0xffffffffu, lane < warp ? wv : 0);
    const int total = __reduce_add_sync(0xffffffffu, wv);
    int pos = base + before + inc - c;
#pragma unroll
    for (int j = 0; j < CR_ITEMS; ++j) {
      if (loc[j] >= 0) {
        if (pos < cap) dst[pos] = loc[j];
        ++pos;
      }
    }
    base += total;
  }
  const int count = base;
  if (threadIdx.x == 0) {
    counts[s] = count;
    // A row list that does not fit its capacity (env-sharded ranks: 1.5 x the mean + 256 rows) would silently bias the
    // gradient: raise the device-side error flag NOW.  The step kernels' optimizer phase sees it (dwopt.cuh) and writes
    // NaN into losses_out, so the overflow surfaces asynchronously in the update's own result; minppo_ctx_check reports it.
    if (count > cap) atomicExch(err_flag, MINPPO_ERR_WORKSPACE);
  }
  for (int j = min(count, cap) + static_cast<int>(threadIdx.x); j < cap; j += CR_THREADS) dst[j] = 0;
}

// pass 0: stats[s] = sum adv ; pass 1: stats[EM+s] = sum (adv - stats[s]/mb)^2   (owned rows)
__global__ void __launch_bounds__(MB_THREADS) adv_stats_kernel(const float* __restrict__ adv,
                                                               const int32_t* __restrict__ rowidx,
                                                               const int32_t* __restrict__ counts,
                                                               float* __restrict__ stats, int EM, int cap,
                                                               int mb, int pass) {
  __shared__ float wsum[MB_THREADS / 32];
  // No early griddep_launch() here: this is the LAST kernel of the per-update chain, and the step kernels that follow
  // read "update-static" data (row lists, statistics) before their own griddepcontrol.wait -- they may only start once
  // this kernel (and, transitively, the whole chain) has completed.
  griddep_wait();
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
                        int cap, int N, int n0, int Nl, int* err_flag, cudaStream_t stream) {
  launch_chain(compact_rows_kernel, dim3(E * M), dim3(CR_THREADS), 0, stream, perms, rowidx, counts, M, B, mb, cap, N, n0, Nl, err_flag);
  return cudaGetLastError() == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
}

int adv_stats_launch(const float* adv, const int32_t* rowidx, const int32_t* counts, float* stats, int EM, int cap,
                     int mb, int pass, cudaStream_t stream) {
  launch_chain(adv_stats_kernel, dim3(EM), dim3(MB_THREADS), 0, stream, adv, rowidx, counts, stats, EM, cap, mb, pass);
  return cudaGetLastError() == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
}

}  // namespace minppo
