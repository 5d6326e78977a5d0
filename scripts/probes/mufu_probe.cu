// Microbenchmark: issue rate of the MUFU (special function) operations the activation epilogues could use, per SM sub-partition.
// One CTA of 512 threads per SM (4 warps per scheduler, as the fused step's worker warps), 8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_probe scripts/probes/mufu_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t y;
  if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  else if (MODE == 1) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
  else if (MODE == 2) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (MODE == 3) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  else if (MODE == 5) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  else asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* cycles, uint32_t* sink) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(0.001f * (threadIdx.x + 37 * i + 1));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = op<MODE>(r[i]);
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= r[i];
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the shape of a tanh epilogue chunk without its memory operations: 16 x (FADD -> MUFU.TANH) -> 8 pack conversions, plus NF
// independent FFMAs standing in for the address arithmetic; reports cycles per chunk and per scheduler (4 warps)
template <int NF, bool WITH_MUFU>
__global__ void __launch_bounds__(512, 1) chunk_probe(int iters, long long* cycles, uint32_t* sink) {
  float v[16], f[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.001f * (threadIdx.x + 37 * i + 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = 1.f + 0.001f * threadIdx.x;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = v[i] + f[i & 3];
      if (WITH_MUFU) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y[i]) : "f"(x));
      else y[i] = fmaxf(x, 0.f);
    }
#pragma unroll
    for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i & 3]) : "f"(1.0001f));
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      uint32_t w;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(y[i + 1]), "f"(y[i]));
      acc ^= w;
      v[i] = __uint_as_float((w << 16) | 0x3a000000u) ; v[i + 1] = __uint_as_float((w & 0xffff0000u) | 0x3a000000u);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = acc + __float_as_uint(f[0] + f[1] + f[2] + f[3]);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int NF, bool WITH_MUFU>
void run_chunk(const char* name, long long* d_cycles, uint32_t* d_sink, int nsm) {
  const int iters = 2048;
  chunk_probe<NF, WITH_MUFU><<<nsm, 512>>>(64, d_cycles, d_sink);
  chunk_probe<NF, WITH_MUFU><<<nsm, 512>>>(iters, d_cycles, d_sink);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d_cycles, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < nsm; ++i) mean += h[i];
  mean /= nsm;
  printf("%-44s %8.1f cycles per chunk (16 elements x 32 lanes) per scheduler\n", name, mean / (double(iters) * 4));
}

template <int MODE>
void run(const char* name, long long* d_cycles, uint32_t* d_sink, int nsm) {
  const int iters = 4096;
  probe<MODE><<<nsm, 512>>>(64, d_cycles, d_sink);
  probe<MODE><<<nsm, 512>>>(iters, d_cycles, d_sink);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d_cycles, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < nsm; ++i) mean += h[i];
  mean /= nsm;
  const double warp_instr_per_smsp = double(iters) * 8 * 4;          // 4 warps per scheduler
  printf("%-24s %8.2f cycles per warp instruction per SMSP  (%5.2f lanes/clk/SMSP)\n", name, mean / warp_instr_per_smsp,
         32.0 * warp_instr_per_smsp / mean);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  long long* d_cycles; uint32_t* d_sink;
  cudaMalloc(&d_cycles, sizeof(long long) * 256);
  cudaMalloc(&d_sink, 16);
  printf("%s, %d SMs, 512 threads per CTA, 8 chains per thread\n", prop.name, nsm);
  run<0>("ex2.approx.ftz.f32", d_cycles, d_sink, nsm);
  run<1>("tanh.approx.f32", d_cycles, d_sink, nsm);
  run<2>("tanh.approx.f16x2", d_cycles, d_sink, nsm);
  run<3>("tanh.approx.bf16x2", d_cycles, d_sink, nsm);
  run<4>("rcp.approx.ftz.f32", d_cycles, d_sink, nsm);
  run<5>("ex2.approx.f16x2", d_cycles, d_sink, nsm);
  run<6>("fma.rn.f32", d_cycles, d_sink, nsm);
  run_chunk<0, false>("chunk: relu, no extra FFMA", d_cycles, d_sink, nsm);
  run_chunk<0, true>("chunk: tanh, no extra FFMA", d_cycles, d_sink, nsm);
  run_chunk<32, false>("chunk: relu + 32 FFMA", d_cycles, d_sink, nsm);
  run_chunk<32, true>("chunk: tanh + 32 FFMA", d_cycles, d_sink, nsm);
  run_chunk<96, true>("chunk: tanh + 96 FFMA", d_cycles, d_sink, nsm);
  return cudaGetLastError() != cudaSuccess;
}
