// minppo_b200 -- GAE reverse scan (replaces /root/reference/minppo/train.py:185-207).
//
// Time-major [T, N] fp32 reward/value, u8 done (the reference's bool, train.py:169/193),
// fp32 last_val[N]; writes advantages and targets = advantages + value (train.py:205).
//
// HBM-bound: 17 algorithmic bytes per transition (4+4+1 read, 4+4 written).  One thread owns
// VEC consecutive envs (VEC=4: float4 / uchar4 accesses, fully coalesced along N) and walks
// t = T-1 .. 0 keeping (gae, next_value) in registers; loads for UNROLL time steps are issued
// before the dependent FMA chain so UNROLL * 36 B per thread are in flight.
//
// When N alone cannot fill the machine the T axis is cut into `chunks` segments.  GAE is the
// linear recurrence gae_t = delta_t + c_t * gae_{t+1}; pass 1 computes every segment's affine
// map (A = prod c_t, Bc = gae at segment start for zero carry-in), pass 2 composes the maps
// of later segments to obtain each segment's carry-in and reruns the segment storing results.
// With chunks == 1 this is exactly the sequential recurrence of the reference.
#include "common.cuh"
#include "minppo_internal.h"

namespace minppo {

template <int VEC> struct VecT;
template <> struct VecT<4> {
  using F = float4; using U = uchar4;
  static __device__ __forceinline__ void ldf(const float* p, float (&o)[4]) { float4 v = __ldcs(reinterpret_cast<const float4*>(p)); o[0]=v.x;o[1]=v.y;o[2]=v.z;o[3]=v.w; }
  static __device__ __forceinline__ void ldu(const uint8_t* p, float (&o)[4]) { uchar4 v = __ldcs(reinterpret_cast<const uchar4*>(p)); o[0]=v.x?0.f:1.f;o[1]=v.y?0.f:1.f;o[2]=v.z?0.f:1.f;o[3]=v.w?0.f:1.f; }
  static __device__ __forceinline__ void stf(float* p, const float (&o)[4]) { __stcs(reinterpret_cast<float4*>(p), make_float4(o[0],o[1],o[2],o[3])); }
};
template <> struct VecT<1> {
  static __device__ __forceinline__ void ldf(const float* p, float (&o)[1]) { o[0] = __ldcs(p); }
  static __device__ __forceinline__ void ldu(const uint8_t* p, float (&o)[1]) { o[0] = __ldcs(p) ? 0.f : 1.f; }
  static __device__ __forceinline__ void stf(float* p, const float (&o)[1]) { __stcs(p, o[0]); }
};

// Walk t in [t0, t1) downwards.  STORE: write adv/tgt.  Returns through gae/nv (carry) and
// aprod (product of c_t) so the same body serves both passes.
template <int VEC, int UNROLL, bool STORE>
__device__ __forceinline__ void gae_segment(const float* __restrict__ reward, const float* __restrict__ value,
                                            const uint8_t* __restrict__ done, float* __restrict__ adv,
                                            float* __restrict__ tgt, size_t N, size_t col, int t0, int t1,
                                            float gamma, float gl, float (&gae)[VEC], float (&nv)[VEC],
                                            float (&aprod)[VEC]) {
  int t = t1 - 1;
  for (; t - (UNROLL - 1) >= t0; t -= UNROLL) {
    float r[UNROLL][VEC], v[UNROLL][VEC], nd[UNROLL][VEC];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t off = static_cast<size_t>(t - u) * N + col;
      VecT<VEC>::ldf(reward + off, r[u]);
      VecT<VEC>::ldf(value + off, v[u]);
      VecT<VEC>::ldu(done + off, nd[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float a[VEC], g[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float delta = r[u][e] + gamma * nv[e] * nd[u][e] - v[u][e];     // train.py:193
        const float c = gl * nd[u][e];
        gae[e] = delta + c * gae[e];                                            // train.py:194
        aprod[e] *= c;
        nv[e] = v[u][e];
        a[e] = gae[e];
        g[e] = gae[e] + v[u][e];                                                // train.py:205
      }
      if (STORE) {
        const size_t off = static_cast<size_t>(t - u) * N + col;
        VecT<VEC>::stf(adv + off, a);
        VecT<VEC>::stf(tgt + off, g);
      }
    }
  }
  for (; t >= t0; --t) {
    float r[VEC], v[VEC], nd[VEC], a[VEC], g[VEC];
    const size_t off = static_cast<size_t>(t) * N + col;
    VecT<VEC>::ldf(reward + off, r);
    VecT<VEC>::ldf(value + off, v);
    VecT<VEC>::ldu(done + off, nd);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float delta = r[e] + gamma * nv[e] * nd[e] - v[e];
      const float c = gl * nd[e];
      gae[e] = delta + c * gae[e];
      aprod[e] *= c;
      nv[e] = v[e];
      a[e] = gae[e];
      g[e] = gae[e] + v[e];
    }
    if (STORE) {
      VecT<VEC>::stf(adv + off, a);
      VecT<VEC>::stf(tgt + off, g);
    }
  }
}

template <int VEC, int UNROLL>
__global__ void __launch_bounds__(256) gae_single_kernel(const float* __restrict__ reward,
                                                         const float* __restrict__ value,
                                                         const uint8_t* __restrict__ done,
                                                         const float* __restrict__ last_val,
                                                         float* __restrict__ adv, float* __restrict__ tgt,
                                                         int T, size_t N, float gamma, float gl) {
  const size_t col = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * VEC;
  if (col >= N) return;
  float gae[VEC], nv[VEC], ap[VEC];
  VecT<VEC>::ldf(last_val + col, nv);
#pragma unroll
  for (int e = 0; e < VEC; ++e) { gae[e] = 0.f; ap[e] = 1.f; }
  gae_segment<VEC, UNROLL, true>(reward, value, done, adv, tgt, N, col, 0, T, gamma, gl, gae, nv, ap);
}

// chunked: blockDim = (cols_per_block, chunks); shared memory holds each segment's (A, Bc).
template <int VEC, int UNROLL>
__global__ void __launch_bounds__(1024) gae_chunked_kernel(const float* __restrict__ reward,
                                                           const float* __restrict__ value,
                                                           const uint8_t* __restrict__ done,
                                                           const float* __restrict__ last_val,
                                                           float* __restrict__ adv, float* __restrict__ tgt,
                                                           int T, size_t N, float gamma, float gl, int seg_len) {
  extern __shared__ float sm[];                        // [chunks][cols][VEC] x 2
  const int cols = blockDim.x, chunks = blockDim.y;
  const int cx = threadIdx.x, ch = threadIdx.y;
  const size_t col = (static_cast<size_t>(blockIdx.x) * cols + cx) * VEC;
  const bool active = col < N;
  const int t0 = ch * seg_len;
  const int t1 = min(T, t0 + seg_len);
  float* sA = sm;
  float* sB = sm + chunks * cols * VEC;
  float gae[VEC], nv[VEC], ap[VEC];
  if (active && t0 < t1) {
    // next_value at the top of the segment is data, not carry: value[t1] or last_val
    if (t1 == T) VecT<VEC>::ldf(last_val + col, nv);
    else VecT<VEC>::ldf(value + static_cast<size_t>(t1) * N + col, nv);
#pragma unroll
    for (int e = 0; e < VEC; ++e) { gae[e] = 0.f; ap[e] = 1.f; }
    float nv0[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) nv0[e] = nv[e];
    gae_segment<VEC, UNROLL, false>(reward, value, done, adv, tgt, N, col, t0, t1, gamma, gl, gae, nv, ap);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      sA[(ch * cols + cx) * VEC + e] = ap[e];
      sB[(ch * cols + cx) * VEC + e] = gae[e];
      nv[e] = nv0[e];
    }
  } else {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      sA[(ch * cols + cx) * VEC + e] = 1.f;
      sB[(ch * cols + cx) * VEC + e] = 0.f;
    }
  }
  __syncthreads();
  if (!(active && t0 < t1)) return;
  // carry-in = gae at t1, composed from the later segments: g = Bc_k + A_k * g  (k = last .. ch+1)
#pragma unroll
  for (int e = 0; e < VEC; ++e) gae[e] = 0.f;
  for (int k = chunks - 1; k > ch; --k) {
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      gae[e] = sB[(k * cols + cx) * VEC + e] + sA[(k * cols + cx) * VEC + e] * gae[e];
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) ap[e] = 1.f;
  gae_segment<VEC, UNROLL, true>(reward, value, done, adv, tgt, N, col, t0, t1, gamma, gl, gae, nv, ap);
}

// Host-side launch plan (exposed for tests through minppo_gae_plan).
void gae_plan(int T, long long N, int sm_count, int* vec, int* chunks, int* seg_len) {
  *vec = (N % 4 == 0) ? 4 : 1;
  const long long cols = N / *vec;
  // enough threads to cover HBM latency: ~1024 resident threads per SM.  Splitting T costs a second read of the inputs
  // (26 instead of 17 bytes per transition), so from ~384 threads per SM on the single pass with 8 time steps of loads
  // in flight per thread wins (measured: N = 262,144 went from 0.61 to the single-pass rate).
  const long long want = static_cast<long long>(sm_count) * 1024;
  int c = 1;
  if (cols < static_cast<long long>(sm_count) * 384)
    while (cols * c < want && c < 32 && T / (c * 2) >= 8) c *= 2;
  *chunks = c;
  *seg_len = (T + c - 1) / c;
}

int gae_launch(const float* reward, const float* value, const uint8_t* done, const float* last_val, float* adv,
               float* tgt, int T, long long N, float gamma, float gl, int sm_count, int force_chunks,
               cudaStream_t stream) {
  int vec, chunks, seg;
  gae_plan(T, N, sm_count, &vec, &chunks, &seg);
  if (vec == 4) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(reward) | reinterpret_cast<uintptr_t>(value) |
                         reinterpret_cast<uintptr_t>(adv) | reinterpret_cast<uintptr_t>(tgt) |
                         reinterpret_cast<uintptr_t>(last_val);
    if ((al & 15u) || (reinterpret_cast<uintptr_t>(done) & 3u)) vec = 1;
  }
  if (force_chunks > 0) {
    chunks = force_chunks;
    seg = (T + chunks - 1) / chunks;
  }
  const long long cols = (N + vec - 1) / vec;
  if (chunks == 1) {
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((cols + threads - 1) / threads);
    if (vec == 4 && cols < static_cast<long long>(sm_count) * 1024)       // few threads: deeper load pipeline per thread
      gae_single_kernel<4, 8><<<blocks, threads, 0, stream>>>(reward, value, done, last_val, adv, tgt, T,
                                                               static_cast<size_t>(N), gamma, gl);
    else if (vec == 4)
      gae_single_kernel<4, 4><<<blocks, threads, 0, stream>>>(reward, value, done, last_val, adv, tgt, T,
                                                               static_cast<size_t>(N), gamma, gl);
    else
      gae_single_kernel<1, 8><<<blocks, threads, 0, stream>>>(reward, value, done, last_val, adv, tgt, T,
                                                               static_cast<size_t>(N), gamma, gl);
  } else {
    const int cx = max(1, 256 / chunks);
    dim3 block(cx, chunks);
    const unsigned blocks = static_cast<unsigned>((cols + cx - 1) / cx);
    const size_t smem = static_cast<size_t>(2) * chunks * cx * vec * sizeof(float);
    if (vec == 4)
      gae_chunked_kernel<4, 4><<<blocks, block, smem, stream>>>(reward, value, done, last_val, adv, tgt, T,
                                                                static_cast<size_t>(N), gamma, gl, seg);
    else
      gae_chunked_kernel<1, 8><<<blocks, block, smem, stream>>>(reward, value, done, last_val, adv, tgt, T,
                                                                static_cast<size_t>(N), gamma, gl, seg);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace minppo
