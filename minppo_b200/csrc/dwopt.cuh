// minppo_b200 -- weight-gradient GEMMs + gradient reduction + global-norm clip + Adam in ONE launch
// (replaces jax.value_and_grad's dW contractions, optax.chain(clip_by_global_norm, adam) and
//  TrainState.apply_gradients: /root/reference/minppo/train.py:98-124, 246-248).
//
// Grid = one CTA per SM, all co-resident (two grid barriers):
//   phase 1  CTAs [0, gemm_ctas): split-K tcgen05 GEMM dW_l = act_l^T dZ_{l+1} for every hidden layer of both
//            nets (umma_gemm_body<EPI_PARTIAL>), fp32 partial tiles -> L2 by TMA store.
//            CTAs [gemm_ctas, grid): meanwhile sum the per-tile partials of the SMALL leaves (biases, output
//            heads, log_std: one partial per 128-row tile of the fused step kernel) and the loss sums -> gflat.
//   barrier
//   phase 2  every thread owns <= EPT arena elements: fixed-order sum of the split-K partials (bitwise
//            reproducible, no atomics), optimizer state fetched alongside, sum of squares -> per-block partial.
//   barrier
//   phase 3  clip scale, Adam, params / mu / nu in place, bf16 weight images for the next step's GEMMs.
// With env-sharded ranks (do_apply == 0) the kernel stops after phase 2 with the local gradient SUM in gflat;
// the all-reduce and opt_kernel (adam.cu, apply only) follow.
#pragma once

#include "opt_common.cuh"
#include "umma_gemm.cuh"

namespace minppo {

struct alignas(64) DwOptParams {
  GemmParams gemm;
  OptArgs opt;
  int gemm_ctas;                 // CTAs [0, gemm_ctas) run the GEMM; the others pre-reduce the small leaves
};

MINPPO_DEVINL bool leaf_is_big(const OptLeaf& L) { return L.img_t != nullptr || L.img_n != nullptr; }   // hidden kernels

// fixed-order sum of `nparts` partials, 16 loads in flight
MINPPO_DEVINL float sum_partials16(const float* __restrict__ src, int nparts, size_t stride) {
  float acc = 0.f;
  int p = 0;
  for (; p + 16 <= nparts; p += 16) {
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = __ldcg(src + static_cast<size_t>(p + u) * stride);
#pragma unroll
    for (int u = 0; u < 16; ++u) acc += v[u];
  }
  for (; p < nparts; ++p) acc += __ldcg(src + static_cast<size_t>(p) * stride);
  return acc;
}

// Job list of the extra CTAs: every element of every small leaf, then the two loss sums.
MINPPO_DEVINL void reduce_small_leaves(const OptArgs& a, int first, int stride) {
  int n_small = 2;
  for (int l = 0; l < a.nleaves; ++l)
    if (!leaf_is_big(a.leaf[l])) n_small += (l + 1 < a.nleaves ? a.leaf[l + 1].offset : a.P) - a.leaf[l].offset;
  for (int j = first; j < n_small; j += stride) {
    int x = j, l = 0;
    for (; l < a.nleaves; ++l) {
      if (leaf_is_big(a.leaf[l])) continue;
      const int n = (l + 1 < a.nleaves ? a.leaf[l + 1].offset : a.P) - a.leaf[l].offset;
      if (x < n) break;
      x -= n;
    }
    if (l < a.nleaves) {
      const OptLeaf& L = a.leaf[l];
      a.gflat[L.offset + x] = sum_partials16(L.grad_src + L.src_offset + x, L.nparts, L.part_stride) + L.grad_bias;
    } else {
      a.gflat[a.P + x] = sum_partials16(a.loss_src + a.loss_src_offset + x, a.loss_nparts, a.loss_part_stride);
    }
  }
}

template <int EPT>
__global__ void __launch_bounds__(GEMM_THREADS, 1) dwopt_kernel(const __grid_constant__ DwOptParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float scratch[32];
  __shared__ float s_bcast[4];
  const OptArgs& a = p.opt;
  const int P = a.P;
  const int G = static_cast<int>(gridDim.x), NT = GEMM_THREADS;
  const int b = static_cast<int>(blockIdx.x), t = static_cast<int>(threadIdx.x);
  const bool has_extra = p.gemm_ctas < G;

  // ---- phase 1 ----------------------------------------------------------------------------------
  if (b < p.gemm_ctas) {
    umma_gemm_body<EPI_PARTIAL>(p.gemm, smem_raw);       // PDL wait / trigger inside (TMA producer warp)
  } else {
    griddep_wait();                                      // the small-leaf partials come from the fused step kernel
    if (t == 0) griddep_launch();
    reduce_small_leaves(a, (b - p.gemm_ctas) * NT + t, (G - p.gemm_ctas) * NT);
  }
  const int count = a.do_apply ? __ldcg(a.count) : 0;   // Adam step count BEFORE this step
  float ent = a.entropy_const;                           // A * (0.5 + 0.5 log 2pi) + sum log|scale|
  if (a.do_apply && b == 0 && t == 0 && a.losses_out) {
    // train.py:240 -- evaluated with the PRE-update log_std (nothing is updated before the second barrier)
    for (int j = 0; j < a.A; ++j) ent += logf(fabsf(expf(__ldcg(a.params + a.off_logstd + j))));
  }
  grid_barrier(a.barrier, a.err_flag);

  // ---- phase 2 ----------------------------------------------------------------------------------
  float g[EPT], pv[EPT], mv[EPT], nv[EPT];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int i = (k * G + b) * NT + t;
    pv[k] = 0.f; mv[k] = 0.f; nv[k] = 0.f;
    if (i < P && a.do_apply) { pv[k] = __ldcg(a.params + i); mv[k] = __ldcg(a.mu + i); nv[k] = __ldcg(a.nu + i); }
  }
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int i = (k * G + b) * NT + t;
    g[k] = 0.f;
    if (i < P) {
      const OptLeaf& L = a.leaf[find_leaf_idx(a, i)];
      if (leaf_is_big(L) || !has_extra) {
        g[k] = sum_partials16(L.grad_src + L.src_offset + (i - L.offset), L.nparts, L.part_stride) + L.grad_bias;
        if (!a.do_apply || a.keep_gflat) a.gflat[i] = g[k];
      } else {
        g[k] = __ldcg(a.gflat + i);
      }
      ss = fmaf(g[k], g[k], ss);
    } else if (i < P + 2 && !has_extra) {
      a.gflat[i] = sum_partials16(a.loss_src + a.loss_src_offset + (i - P), a.loss_nparts, a.loss_part_stride);
    }
  }
  if (!a.do_apply) return;

  const float bs = block_sum<GEMM_THREADS>(ss, scratch);
  if (t == 0) a.block_ss[b] = bs;
  grid_barrier(a.barrier, a.err_flag);

  // ---- phase 3 ----------------------------------------------------------------------------------
  if (t < 32) {
    float s = 0.f;
    for (int x = t; x < G; x += 32) s += __ldcg(a.block_ss + x);
    s = warp_sum(s);
    if (t == 0) s_bcast[0] = sqrtf(s);
  }
  if (t == 32) step_scalars(a, count, s_bcast[1], s_bcast[2], s_bcast[3]);
  __syncthreads();
  AdamScalars sc;
  sc.gnorm = s_bcast[0]; sc.lr = s_bcast[1]; sc.c1 = s_bcast[2]; sc.c2 = s_bcast[3];
  sc.trigger = sc.gnorm < a.max_norm;                    // optax.clip_by_global_norm
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int i = (k * G + b) * NT + t;
    if (i >= P) continue;
    adam_element(a, sc, g[k], pv[k], mv[k], nv[k]);
    a.params[i] = pv[k];
    a.mu[i] = mv[k];
    a.nu[i] = nv[k];
    write_images(a.leaf[find_leaf_idx(a, i)], i, pv[k]);
  }
  if (b == 0 && t == 0) {
    *a.count = count + 1;
    if (a.losses_out) {
      // gflat[P] = sum max(vl, vlc), gflat[P+1] = sum min(l1, l2) over the global minibatch
      const float value_loss = 0.5f * __ldcg(a.gflat + P) * a.inv_mb;
      const float actor_loss = -__ldcg(a.gflat + P + 1) * a.inv_mb;
      a.losses_out[0] = actor_loss + a.vf_coef * value_loss - a.ent_coef * ent;
      a.losses_out[1] = value_loss;
      a.losses_out[2] = actor_loss;
      a.losses_out[3] = ent;
      if (a.gnorm_out) *a.gnorm_out = sc.gnorm;
    }
  }
}

constexpr int DWOPT_MAX_EPT = 12;
inline long long dwopt_max_params(int grid) { return static_cast<long long>(grid) * GEMM_THREADS * DWOPT_MAX_EPT - 2; }

inline cudaError_t dwopt_launch(const DwOptParams& p, int grid, cudaStream_t stream, bool pdl) {
  const long long per_thread = (static_cast<long long>(p.opt.P) + 2 + static_cast<long long>(grid) * GEMM_THREADS - 1) /
                               (static_cast<long long>(grid) * GEMM_THREADS);
  if (per_thread <= 4) return launch_kernel(dwopt_kernel<4>, grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream, pdl, p);
  if (per_thread <= 8) return launch_kernel(dwopt_kernel<8>, grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream, pdl, p);
  return launch_kernel(dwopt_kernel<DWOPT_MAX_EPT>, grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream, pdl, p);
}

}  // namespace minppo
