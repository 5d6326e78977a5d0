// minppo_b200 -- gradient reduction + global-norm clip + Adam, one multi-tensor kernel
// (replaces optax.chain(clip_by_global_norm, adam) and TrainState.apply_gradients,
//  /root/reference/minppo/train.py:98-101, 114-124, 248).
//
// The whole parameter set lives in one contiguous fp32 arena (leaf order = JAX's sorted
// flatten order of the pickle tree); mu / nu are arenas of the same shape.
//
// Phase R (reduce): every gradient element is the fixed-order sum of the per-tile / per-split
//   partials the backward kernels wrote (no atomics -> bitwise reproducible), written to gflat;
//   the minibatch loss sums ride along at gflat[P .. P+4).
// Phase A (apply): sum of squares -> grid barrier -> clip scale -> Adam -> params, mu, nu and the
//   bf16 weight images the tcgen05 GEMMs read (transposed [out][in] for forward, [in][out] for dX).
// Single GPU runs R+A in one launch; with env-sharded ranks R, NCCL all-reduce(gflat), A.
#include "common.cuh"
#include "minppo_internal.h"
#include "opt_common.cuh"

namespace minppo {

// EPT elements per thread, kept in registers across the grid barrier.  Elements are dealt to
// warps in 32-element chunks, round-robin over BLOCKS (chunk = k*G*32 + warp*G + block), so that
// the few leaves with many partials (output heads) are spread over the whole grid while every
// warp still reads 128 contiguous bytes per partial.
template <int EPT>
__global__ void __launch_bounds__(OPT_THREADS, 3) opt_kernel(const OptArgs a) {
  __shared__ float scratch[32];
  __shared__ float s_bcast[4];
  const int P = a.P;
  const int G = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int count = a.do_apply ? __ldcg(a.count) : 0;            // Adam step count BEFORE this step
  float ent = a.entropy_const;                             // A * (0.5 + 0.5 log 2pi) + sum log|scale|
  if (a.do_apply && blockIdx.x == 0 && threadIdx.x == 0 && a.losses_out) {
    // train.py:240 -- evaluated with the PRE-update log_std (nothing is updated before the barrier)
    for (int j = 0; j < a.A; ++j) ent += logf(fabsf(expf(__ldcg(a.params + a.off_logstd + j))));
  }

  float g[EPT], pv[EPT], mv[EPT], nv[EPT];
  float ss = 0.f;
  // optimizer state of this thread's elements: independent of the gradient producers, so it is
  // fetched before the PDL wait (nothing else writes params / mu / nu between two optimizer steps)
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int i = ((k * (OPT_THREADS / 32) + warp) * G + static_cast<int>(blockIdx.x)) * 32 + lane;
    pv[k] = 0.f; mv[k] = 0.f; nv[k] = 0.f;
    if (i < P && a.do_apply) { pv[k] = __ldcg(a.params + i); mv[k] = __ldcg(a.mu + i); nv[k] = __ldcg(a.nu + i); }
  }
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch();
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int i = ((k * (OPT_THREADS / 32) + warp) * G + static_cast<int>(blockIdx.x)) * 32 + lane;
    g[k] = 0.f;
    if (i < P) {
      if (a.do_reduce) {
        const OptLeaf& L = a.leaf[find_leaf_idx(a, i)];
        g[k] = sum_partials(L.grad_src + L.src_offset + (i - L.offset), L.nparts, L.part_stride) + L.grad_bias;
        if (!a.do_apply || a.keep_gflat) a.gflat[i] = g[k];
      } else {
        g[k] = __ldcg(a.gflat + i);
      }
      if (a.do_apply) ss = fmaf(g[k], g[k], ss);
    } else if (i < P + 2 && a.do_reduce) {
      a.gflat[i] = sum_partials(a.loss_src + a.loss_src_offset + (i - P), a.loss_nparts, a.loss_part_stride);
    }
  }
  if (!a.do_apply) return;

  const float bs = block_sum(ss, scratch);
  if (threadIdx.x == 0) a.block_ss[blockIdx.x] = bs;
  grid_barrier(a.barrier, a.err_flag);
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int b = threadIdx.x; b < G; b += 32) s += __ldcg(a.block_ss + b);
    s = warp_sum(s);
    if (threadIdx.x == 0) s_bcast[0] = sqrtf(s);
  }
  if (threadIdx.x == 32) {                                 // per-step scalars, once per block
    float lr;                                              // train.py:98-101 (annealed) or opt.lr
    if (a.anneal) {
      const float frac = 1.0f - static_cast<float>(count / a.anneal_div) / static_cast<float>(a.num_updates);
      lr = a.lr * frac;
    } else {
      lr = a.lr;
    }
    const float cnt1 = static_cast<float>(count + 1);
    s_bcast[1] = lr;
    s_bcast[2] = 1.0f - powf(a.b1, cnt1);
    s_bcast[3] = 1.0f - powf(a.b2, cnt1);
  }
  __syncthreads();
  const float gnorm = s_bcast[0];
  const bool trigger = gnorm < a.max_norm;                 // optax.clip_by_global_norm
  const float lr = s_bcast[1], c1 = s_bcast[2], c2 = s_bcast[3];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int i = ((k * (OPT_THREADS / 32) + warp) * G + static_cast<int>(blockIdx.x)) * 32 + lane;
    if (i >= P) continue;
    float gg = g[k];
    if (!trigger) gg = (gg / gnorm) * a.max_norm;
    const float mu = a.one_minus_b1 * gg + a.b1 * mv[k];
    const float nu = a.one_minus_b2 * (gg * gg) + a.b2 * nv[k];
    const float u = (mu / c1) / (sqrtf(nu / c2 + a.eps_root) + a.eps);
    const float p = pv[k] + (-lr) * u;
    a.params[i] = p;
    a.mu[i] = mu;
    a.nu[i] = nu;
    write_images(a.leaf[find_leaf_idx(a, i)], i, p);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *a.count = count + 1;
    if (a.losses_out) {
      // gflat[P] = sum max(vl, vlc), gflat[P+1] = sum min(l1, l2) over the global minibatch
      const float value_loss = 0.5f * __ldcg(a.gflat + P) * a.inv_mb;
      const float actor_loss = -__ldcg(a.gflat + P + 1) * a.inv_mb;
      a.losses_out[0] = actor_loss + a.vf_coef * value_loss - a.ent_coef * ent;
      a.losses_out[1] = value_loss;
      a.losses_out[2] = actor_loss;
      a.losses_out[3] = ent;
      if (a.gnorm_out) *a.gnorm_out = gnorm;
    }
  }
}

// fp32 [rows][cols] -> bf16 [rows][ld] (zero padded); used for the observation image and
// for the initial weight images.
__global__ void f32_to_bf16_image_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                         long long rows, int cols, int ld) {
  const long long pairs_per_row = ld / 2;
  const long long total = rows * pairs_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / pairs_per_row;
    const int c = static_cast<int>(i % pairs_per_row) * 2;
    const float* s = src + r * cols;
    const float x0 = c < cols ? s[c] : 0.f;
    const float x1 = c + 1 < cols ? s[c + 1] : 0.f;
    reinterpret_cast<uint32_t*>(dst)[i] = pack_bf16x2(x0, x1);
  }
}

// weight images from the fp32 arena (ctx_create / set_params): same mapping as the apply phase
__global__ void weight_images_kernel(const OptArgs a) {
  const int gthreads = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.P; i += gthreads) {
    int l = 0;
    while (l + 1 < a.nleaves && i >= a.leaf[l + 1].offset) ++l;
    write_images(a.leaf[l], i, a.params[i]);
  }
}

int opt_max_params(int blocks) { return blocks * OPT_THREADS * OPT_EPT - 2; }

int opt_launch(const OptArgs& a, int blocks, cudaStream_t stream, bool pdl) {
  if (a.P > opt_max_params(blocks)) return MINPPO_ERR_UNSUPPORTED;   // single sweep: global norm needs all elements
  const long long per_thread = (static_cast<long long>(a.P) + 2 + static_cast<long long>(blocks) * OPT_THREADS - 1) /
                               (static_cast<long long>(blocks) * OPT_THREADS);
  cudaError_t e;
  if (per_thread <= 2) e = launch_kernel(opt_kernel<2>, blocks, OPT_THREADS, 0, stream, pdl, a);
  else if (per_thread <= 4) e = launch_kernel(opt_kernel<4>, blocks, OPT_THREADS, 0, stream, pdl, a);
  else e = launch_kernel(opt_kernel<OPT_EPT>, blocks, OPT_THREADS, 0, stream, pdl, a);
  return e == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
}

int weight_images_launch(const OptArgs& a, cudaStream_t stream) {
  weight_images_kernel<<<148, 256, 0, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
}

int obs_image_launch(const float* obs, __nv_bfloat16* img, long long rows, int cols, int ld, cudaStream_t stream) {
  const long long total = rows * (ld / 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (blocks < 1) blocks = 1;
  f32_to_bf16_image_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(obs, img, rows, cols, ld);
  return cudaGetLastError() == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
}

}  // namespace minppo
