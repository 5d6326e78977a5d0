// Microbenchmark: the fused step's activation / derivative epilogues in isolation (no MMA, no TMA traffic): 16 worker warps
// walk a 128 x 256 fp32 accumulator in TMEM exactly as epilogues 1 (column stride 64) and 2 (stride 16) do.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I minppo_b200/csrc -o scripts/probes/epilogue_probe scripts/probes/epilogue_probe.cu -lcuda
#include <cstdio>
#include "fused_step.cuh"

using namespace minppo;

template <int KIND>   // 0: act stride 16, 1: act stride 64, 2: dact stride 64, 3: dact stride 64 publishing every chunk on an mbarrier
__global__ void __launch_bounds__(FS_THREADS, 1) probe(int act, int iters, int sync_each, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ uint64_t pub[4];                                   // never waited on: arrival counts far above what the probe delivers
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&pub[i], (1u << 20) - 1u); fence_mbar_init(); }
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  float* bias_s = reinterpret_cast<float*>(smem_raw + (base - raw) + 131072);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == FS_MMA_WARP) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  if (threadIdx.x < 512) bias_s[threadIdx.x] = 0.001f * threadIdx.x;
  for (int i = threadIdx.x; i < 32768; i += FS_THREADS) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0x3c003c00u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = slot;
  if (warp < 16) {
    const int q = warp & 3, sub = warp >> 2, erow = q * 32 + lane;
    worker_bar();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (sync_each) worker_bar();                               // single-shot: every pass starts with all 16 warps aligned, as after an mbarrier
      if (KIND == 0) epilogue_act(acc, base, bias_s, act, erow, q, sub * 64, 16, 4, nullptr);
      else if (KIND == 1) epilogue_act(acc, base, bias_s, act, erow, q, sub * 16, 64, 4, nullptr);
      else if (KIND == 2) epilogue_dact(acc, base + 65536, base, act, erow, q, sub * 16, 64, 4, nullptr);
      else epilogue_dact(acc, base + 65536, base, act, erow, q, sub * 16, 64, 4, pub);
    }
    worker_bar();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FS_MMA_WARP) tmem_dealloc(acc, 512);
}

template <int KIND>
void run(const char* name, int act, long long* d_cycles, int nsm) {
 for (int sync_each = 0; sync_each < 2; ++sync_each) {
  const int iters = 200, smem = 131072 + 2048 + 1024;
  cudaFuncSetAttribute(probe<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<KIND><<<nsm, FS_THREADS, smem>>>(act, 4, sync_each, d_cycles);
  probe<KIND><<<nsm, FS_THREADS, smem>>>(act, iters, sync_each, d_cycles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, d_cycles, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < nsm; ++i) mean += h[i];
  printf("%-52s %8.1f cycles per 128 x 256 tile (%s)\n", name, mean / nsm / iters, sync_each ? "single shot: barrier before every pass" : "back to back");
 }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 256);
  printf("%s, %d SMs, 16 epilogue warps per CTA\n", prop.name, nsm);
  run<0>("H2 epilogue (stride 16), relu", ACT_RELU, d_cycles, nsm);
  run<0>("H2 epilogue (stride 16), tanh.approx", ACT_TANH_FAST, d_cycles, nsm);
  run<0>("H2 epilogue (stride 16), tanh via ex2 + rcp", ACT_TANH, d_cycles, nsm);
  run<1>("H1 epilogue (stride 64), relu", ACT_RELU, d_cycles, nsm);
  run<1>("H1 epilogue (stride 64), tanh.approx", ACT_TANH_FAST, d_cycles, nsm);
  run<2>("dZ epilogue (stride 64), relu", ACT_RELU, d_cycles, nsm);
  run<2>("dZ epilogue (stride 64), tanh", ACT_TANH, d_cycles, nsm);
  run<3>("dZ epilogue, published per chunk, relu", ACT_RELU, d_cycles, nsm);
  run<3>("dZ epilogue, published per chunk, tanh", ACT_TANH, d_cycles, nsm);
  return 0;
}
