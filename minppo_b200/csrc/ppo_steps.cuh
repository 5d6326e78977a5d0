// minppo_b200 -- ALL minibatch steps of one learner update in ONE persistent launch
// (the epoch / minibatch scans of /root/reference/minppo/train.py:268, 274 with everything inside them: 213-249).
//
// Grid = one CTA per SM, all co-resident, FS_THREADS threads; the 512 TMEM columns are allocated once.  Per step s:
//
//   phase A   CTA b runs the (tile, net) units b, b + grid, ... of fused_tile (forward + loss + backward-to-dZ)
//   grid barrier          H1 / dZ / X tiles and the per-tile partials are visible everywhere
//   phase B   dwopt_body: split-K weight-gradient GEMM -> barrier -> fixed-order reduction (+ the NVLink peer exchange
//             on env-sharded ranks) -> barrier -> global-norm clip + Adam + bf16 weight images
//   grid barrier          the new parameters / images are visible before step s + 1 reads them
//
// Against the two launches per step this removes 2 x E x M kernel boundaries (the dependent-launch latency, both
// prologues -- TMEM allocation, tensor-map and constant-bank fetches -- and the completion flush of every boundary; on a
// context with peer mappings each boundary costs ~1.3 us more), at the price of two more grid barriers per step.
// Strictly sequential SGD is preserved: every phase of step s completes (grid-wide) before the next one starts.
#pragma once

#include "dwopt.cuh"
#include "fused_step.cuh"

namespace minppo {

struct alignas(64) StepsParams {
  FusedParams fs;
  DwOptParams dw;
  int s0, s1;                    // minibatch steps [s0, s1) of this launch
  int units;                     // (tile, net) units per step: 2 * m_tiles
};

template <int AP>
constexpr int steps_smem_bytes() { return FsLayout<AP>::BYTES > GEMM_SMEM_BYTES ? FsLayout<AP>::BYTES : GEMM_SMEM_BYTES; }

template <int AP, int MAXU>
__global__ void __launch_bounds__(FS_THREADS, 1) ppo_steps_kernel(const __grid_constant__ StepsParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5;
  const int b = static_cast<int>(blockIdx.x), G = static_cast<int>(gridDim.x);
  if (warp == FS_MMA_WARP) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // debug stamps (MINPPO_TRACE): row b of the fused kernel's trace buffer, slots the tile body leaves free
  long long* tr = (P.fs.trace && b < P.units && threadIdx.x == 0) ? P.fs.trace + static_cast<size_t>(b) * FS_TRACE_SLOTS : nullptr;
#define PS_STAMP(slot) do { if (tr) tr[(slot)] = clock64(); } while (0)
  for (int s = P.s0; s < P.s1; ++s) {
    PS_STAMP(13);
    for (int u = b; u < P.units; u += G) fused_tile<AP, true>(P.fs, u, s, sm, base, tmem_base);
    PS_STAMP(14);
    grid_barrier(P.dw.opt.barrier, P.dw.opt.err_flag);
    PS_STAMP(15);
    dwopt_body<MAXU>(P.dw, s, smem_raw, tmem_base, s == P.s0);
    PS_STAMP(20);
    grid_barrier(P.dw.opt.barrier, P.dw.opt.err_flag);
    PS_STAMP(22);
  }
#undef PS_STAMP

  tc_fence_before();
  __syncthreads();
  if (warp == FS_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

template <int AP, int MAXU>
inline cudaError_t steps_launch_t(const StepsParams& p, int grid, cudaStream_t stream) {
  return launch_kernel(ppo_steps_kernel<AP, MAXU>, grid, FS_THREADS, steps_smem_bytes<AP>(), stream, false, p);
}
inline cudaError_t steps_launch(const StepsParams& p, int grid, cudaStream_t stream, int ap, int maxu) {
  if (ap == 16) {
    if (maxu <= 1) return steps_launch_t<16, 1>(p, grid, stream);
    if (maxu == 2) return steps_launch_t<16, 2>(p, grid, stream);
    return steps_launch_t<16, 4>(p, grid, stream);
  }
  if (maxu <= 1) return steps_launch_t<32, 1>(p, grid, stream);
  if (maxu == 2) return steps_launch_t<32, 2>(p, grid, stream);
  return steps_launch_t<32, 4>(p, grid, stream);
}
template <int AP, int MAXU>
inline cudaError_t steps_attr_t() {
  return cudaFuncSetAttribute(ppo_steps_kernel<AP, MAXU>, cudaFuncAttributeMaxDynamicSharedMemorySize, steps_smem_bytes<AP>());
}
inline cudaError_t steps_init_attrs() {
  cudaError_t e = steps_attr_t<16, 1>();
  if (e == cudaSuccess) e = steps_attr_t<16, 2>();
  if (e == cudaSuccess) e = steps_attr_t<16, 4>();
  if (e == cudaSuccess) e = steps_attr_t<32, 1>();
  if (e == cudaSuccess) e = steps_attr_t<32, 2>();
  if (e == cudaSuccess) e = steps_attr_t<32, 4>();
  return e;
}

}  // namespace minppo
