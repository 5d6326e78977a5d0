// Microbenchmark: TMEM -> register read throughput per SM for the tcgen05.ld shapes / widths the epilogues could use.
// One CTA per SM, NW warps; every warp reads ITER x (all 512 columns of its lane quadrant) and keeps a checksum.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_probe scripts/probes/tmem_ld_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__device__ __forceinline__ uint32_t ld_chunk(uint32_t taddr) {
  uint32_t acc = 0;
  if (MODE == 0) {          // 32x32b.x16: 32 lanes x 16 columns = 2 KB per warp instruction
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else if (MODE == 1) {   // 32x32b.x32: 4 KB
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),"=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  } else if (MODE == 2) {   // 16x256b.x4: 16 lanes x 256 bit x 4 = 32 columns of 16 lanes -> 2 KB, two instructions cover 32 lanes
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else if (MODE == 3) {   // two 32x32b.x16 in flight before one wait
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(taddr + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  }
  return acc;
}
template <int MODE> __device__ constexpr int chunk_cols() { return MODE == 0 ? 16 : 32; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  constexpr int CC = chunk_cols<MODE>();
  // the warps sharing a lane quadrant split the 512 columns between them (like epilogue warps do)
  const int nshare = blockDim.x >> 7, share = warp >> 2;
  for (int it = 0; it < iters; ++it)
    for (int c = share * CC; c < 512; c += nshare * CC) {
      if (MODE == 2) { acc ^= ld_chunk<2>(base + c); acc ^= ld_chunk<2>(base + c + (16u << 16)); }
      else acc ^= ld_chunk<MODE>(base + c);
    }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
void run(const char* name, int nwarps) {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 64;
  probe<MODE><<<148, nwarps * 32>>>(4, cyc, sink);
  probe<MODE><<<148, nwarps * 32>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
  const double bytes = (double)iters * 128 * 512 * 4;            // the whole 256 KB of TMEM, iters times, per SM
  printf("%-34s warps %2d: %9.0f cycles  -> %6.1f B/clk/SM  (%s)\n", name, nwarps, mean, bytes / mean, cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int nw : {4, 8, 16, 32}) {
    run<0>("32x32b.x16 + wait", nw);
    run<1>("32x32b.x32 + wait", nw);
    run<3>("2 x 32x32b.x16, one wait", nw);
    run<2>("16x256b.x4 (two per 32 lanes) + wait", nw);
  }
  return 0;
}
