// minppo_b200 -- fused minibatch forward + PPO loss + backward-to-dZ for the 2-hidden-layer nets
// (the reference's default model.num_layers = 2, /root/reference/minppo/config.py:53).
//
// One CTA = one 128-row tile of the minibatch x one net (actor or critic).  Everything between
// the gathered observation rows and the layer gradients dZ stays on chip; every contraction runs
// on the tcgen05 tensor cores with accumulators in TMEM:
//
//   gather X (cp.async by row index, SW128)               -> smem R1
//   L1   acc0 = X  W0^T  (B streamed by TMA)              -> +b0, act, bf16 -> R0 (H1) -> TMA store
//   L2   acc1 = H1 W1^T                                   -> +b1, act, bf16 -> R1 (H2)
//   head out  = H2 W2    (N = 16)                         -> +b2 -> Gaussian log-prob / clipped
//        surrogate (actor CTA) or clipped value loss (critic CTA), train.py:218-243 -> g = dL/dout
//   bwd  dA2  = g W2^T   (K = 16),  dW2 = H2^T g (N = 16) -> dZ2 = dA2 * f'(H2) in place -> TMA store
//   dH1  acc1 = dZ2 W1                                    -> * f'(H1), bf16 -> R1 (dZ1) -> TMA store
//
// The output-head operands W2 and g are fp32 quantities: they enter the tensor cores as a bf16
// hi/lo pair (x = hi + lo to 2^-17), the cross terms hi*hi + hi*lo + lo*hi are accumulated in
// fp32, so the heads keep fp32-level accuracy while costing a few dozen tiny UMMAs.
//
// Warp roles: warps 0..7 workers (gather, epilogues, loss), two per TMEM lane quadrant; warp 8 TMA
// producer (weight k-blocks through a 2-stage ring); warp 9 TMEM allocator + MMA issuer.  H1, dZ2 and dZ1 are also written to HBM (bf16, TMA stores) for the split-K
// weight-gradient GEMM (umma_gemm.cuh, EPI_PARTIAL), which runs as its own launch over all tiles.
#pragma once

#include "common.cuh"
#include "minppo_internal.h"
#include "umma_gemm.cuh"

namespace minppo {

constexpr int FS_THREADS = 320;
constexpr int FS_WORKERS = 256;                         // warps 0..7
constexpr int FS_TMA_WARP = 8;                          // weight producer
constexpr int FS_MMA_WARP = 9;                          // TMEM allocator + MMA issuer: the highest warp id wins the
                                                        // issue arbiter, so the single issuing thread is never starved
constexpr int FS_AP = 16;                               // padded head width (A <= 16)
constexpr int FS_R0 = 0;                                // H1                  64 KB
constexpr int FS_R1 = 65536;                            // X / H2 / dZ2 / dZ1  64 KB
constexpr int FS_RB = 131072;                           // weight ring   2 x 32 KB
constexpr int FS_BSTAGE = 32768;
constexpr int FS_W2T = 196608;                          // head kernel^T bf16 [32][H] SW128: rows 0..15 hi, 16..31 lo; 4 KB per 64 columns
constexpr int FS_GT = FS_W2T + 16384;                   // g^T bf16 [32][128] SW128: rows 0..15 hi, 16..31 lo; 4 KB per 64 rows
constexpr int FS_BIAS = FS_GT + 8192;                   // [2][256] f32
constexpr int FS_HB = FS_BIAS + 2048;                   // f32: head bias [16], log_std [16], 1/scale [16], log-det [1]
constexpr int FS_RED = FS_HB + 256;                     // [8][40] f32
constexpr int FS_BARS = FS_RED + 8 * 40 * 4;            // mbarriers + tmem slot
constexpr int FS_SMEM_BYTES = FS_BARS + 256 + 1024;     // + alignment slack

struct alignas(64) FusedNet {
  CUtensorMap tm_w0;             // W0 image [Dp][H] (as stored: [in][out])  box {64 out, 64 in}: MN-major B of L1
  CUtensorMap tm_w1;             // W1 image [H][H]                          box {64 out, 64 in}: MN-major B of L2
  CUtensorMap tm_w1k;            // W1 image [H][H]                          box {64 out, H in}:  K-major  B of dH1 (n = in, k = out)
  CUtensorMap tm_h1;             // act[1] [M_pad][H]    box {64, 128}  (TMA store)
  CUtensorMap tm_dz2;            // dz[2]
  CUtensorMap tm_dz1;            // dz[1]
  const float* b0;               // arena pointers
  const float* b1;
  const uint4* w2img;            // head kernel^T as bf16 hi / lo, the 16 KB smem image of FS_W2T (written by the optimizer)
  const float* b2;               // head bias [aout]
  int act;                       // ACT_*
  int aout;                      // A (actor) or 1 (critic)
  int po_w2, po_b2, po_loss;     // offsets of this net's fields inside a head partial
};

struct alignas(64) FusedParams {
  FusedNet net[2];
  CUtensorMap tm_xg;             // gathered observation rows [M_pad][Dp], box {64, 128}: TMA store by the actor CTAs,
                                 // the A operand of both nets' first-layer weight-gradient GEMM
  const int32_t* rowidx;         // [cap] (this minibatch)
  const __nv_bfloat16* obs_img;  // [Bl][Dp]
  const int32_t* count;
  const float* adv_sum;
  const float* adv_sq;
  const float* action;
  const float* v_old;
  const float* logp_old;
  const float* adv;
  const float* tgt;
  const float* log_std;          // arena pointer
  float* part;                   // head partials [m_tiles][part_stride]
  int part_stride, po_logstd;
  int H, A, Dp, m_tiles, cap;
  float inv_mb, clip_eps, vf_coef;
  long long* trace;              // debug: [ctas][32] clock64 stamps (null = off)
};

#define FS_STAMP(slot) do { if (p.trace) p.trace[static_cast<size_t>(cta_id) * 32 + (slot)] = clock64(); } while (0)

MINPPO_DEVINL void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

MINPPO_DEVINL void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)) : "memory");
}
// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns
MINPPO_DEVINL void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// byte offset of the 16-byte chunk holding column c of row r inside a [128][H] bf16 SW128 tile set
MINPPO_DEVINL uint32_t sw_off(int r, int c) {
  return static_cast<uint32_t>((c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4));
}
template <int ACT>
MINPPO_DEVINL float act_apply(float x) {
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == ACT_TANH_FAST) return fast_tanh(x);
  return exp_tanh(x);
}
template <int ACT>
MINPPO_DEVINL float act_deriv_t(float h) { return ACT == ACT_RELU ? (h > 0.f ? 1.f : 0.f) : (1.f - h * h); }
MINPPO_DEVINL float dclip_f(float x, float lo, float hi) {
  return (x > lo && x < hi) ? 1.f : ((x == lo || x == hi) ? 0.5f : 0.f);
}
// accumulator (32 columns per chunk) -> +bias, activation, bf16 -> swizzled smem tile
// (not inlined: epilogue 1 and 2 share one copy of the code -- the kernel's straight-line worker path is larger than
//  the instruction cache, and "no instruction" was a quarter of its stall samples)
template <int ACT>
__device__ __noinline__ void epilogue_act_t(uint32_t tmem_acc, uint32_t dst_base, const float* bias_s, int row, int q, int col0,
                                  int ncols) {
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  for (int c0 = col0; c0 < col0 + ncols; c0 += 32) {
    float v[32];
    tmem_ld_32x32(taddr + c0, v);
    const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);     // broadcast LDS.128
    float4 bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bb[j] = b4[j];
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t w[4];
      const float bj[8] = {bb[2 * j].x, bb[2 * j].y, bb[2 * j].z, bb[2 * j].w,
                           bb[2 * j + 1].x, bb[2 * j + 1].y, bb[2 * j + 1].z, bb[2 * j + 1].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int e = 8 * j + 2 * t;
        const float x0 = act_apply<ACT>(v[e] + bj[2 * t]);
        const float x1 = act_apply<ACT>(v[e + 1] + bj[2 * t + 1]);
        w[t] = pack_bf16x2(x0, x1);
      }
      sts128(dst_base + sw_off(row, c0 + 8 * j), make_uint4(w[0], w[1], w[2], w[3]));
    }
  }
}
MINPPO_DEVINL void epilogue_act(uint32_t tmem_acc, uint32_t dst_base, const float* bias_s, int act, int row, int q,
                                int col0, int ncols) {
  if (act == ACT_RELU) epilogue_act_t<ACT_RELU>(tmem_acc, dst_base, bias_s, row, q, col0, ncols);
  else if (act == ACT_TANH_FAST) epilogue_act_t<ACT_TANH_FAST>(tmem_acc, dst_base, bias_s, row, q, col0, ncols);
  else epilogue_act_t<ACT_TANH>(tmem_acc, dst_base, bias_s, row, q, col0, ncols);
}

// dZ = acc * f'(h): h read from `h_base`, bf16 result written to `dst_base` (may alias h_base:
// every thread touches only its own 16-byte chunks).  The bias gradients (column sums of dZ) are
// not formed here: the weight-gradient GEMM gets them from the tensor core as ones x dZ.
template <int ACT>
__device__ __noinline__ void epilogue_dact_t(uint32_t tmem_acc, uint32_t h_base, uint32_t dst_base, int row, int q, int col0,
                                   int ncols) {
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  for (int c0 = col0; c0 < col0 + ncols; c0 += 32) {
    float v[32];
    tmem_ld_32x32(taddr + c0, v);
    uint4 hh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) hh[j] = lds128(h_base + sw_off(row, c0 + 8 * j));
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t hw[4] = {hh[j].x, hh[j].y, hh[j].z, hh[j].w};
      uint32_t w[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int e = 8 * j + 2 * t;
        w[t] = pack_bf16x2(v[e] * act_deriv_t<ACT>(bf16_lo(hw[t])), v[e + 1] * act_deriv_t<ACT>(bf16_hi(hw[t])));
      }
      sts128(dst_base + sw_off(row, c0 + 8 * j), make_uint4(w[0], w[1], w[2], w[3]));
    }
  }
}
MINPPO_DEVINL void epilogue_dact(uint32_t tmem_acc, uint32_t h_base, uint32_t dst_base, int act, int row, int q,
                                 int col0, int ncols) {
  if (act == ACT_RELU) epilogue_dact_t<ACT_RELU>(tmem_acc, h_base, dst_base, row, q, col0, ncols);
  else epilogue_dact_t<ACT_TANH>(tmem_acc, h_base, dst_base, row, q, col0, ncols);
}

__global__ void __launch_bounds__(FS_THREADS, 1) fused_step_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* bias_s = reinterpret_cast<float*>(sm + FS_BIAS);       // [0..256) layer 0, [256..512) layer 1
  float* hb = reinterpret_cast<float*>(sm + FS_HB);             // [0..16) head bias, [16..32) log_std,
                                                                // [32..48) 1 / scale, [48] sum log|scale|
  float* red = reinterpret_cast<float*>(sm + FS_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + FS_BARS);
  uint64_t* full_bar = bars;            // [2]
  uint64_t* empty_bar = bars + 2;       // [2]
  uint64_t* accf0 = bars + 4;           // L1 accumulator complete
  uint64_t* accf1 = bars + 5;           // L2 accumulator complete
  uint64_t* headf = bars + 6;           // head outputs complete
  uint64_t* bwdf = bars + 7;            // dA2 and dW2 complete
  uint64_t* dh1f = bars + 8;            // dH1 accumulator complete
  uint64_t* xfull = bars + 9;           // workers -> MMA: X tile gathered
  uint64_t* h2r = bars + 10;            // H2 in R1, acc1 drained
  uint64_t* gr = bars + 11;             // g^T hi/lo written
  uint64_t* h1r = bars + 12;            // [4] H1 columns [64 b, 64 b + 64) in R0 (k-block b of the L2 GEMM)
  uint64_t* dz2r = bars + 16;           // [4] dZ2 columns [64 b, 64 b + 64) in R1 (k-block b of the dH1 GEMM)
  uint64_t* w0x = bars + 20;            // [2] W0 k-blocks 2, 3 parked in R0 (free until epilogue 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA order: (tile, net) with net fastest, so that the LIVE tiles of both nets are the lowest block indices and fit the
  // first wave even when the grid is sized for an env-sharded rank's worst-case row count (dead tiles come last)
  const int net = static_cast<int>(blockIdx.x) & 1;
  const int tile = static_cast<int>(blockIdx.x) >> 1;
  const int cta_id = net * p.m_tiles + tile;             // trace row (net-major, scripts/trace_fused.py)
  const FusedNet& G = p.net[net];
  const int H = p.H, nkH = H >> 6, nk0 = p.Dp >> 6;
  const uint32_t R0 = base + FS_R0, R1 = base + FS_R1, RB = base + FS_RB;
  const uint32_t W2T = base + FS_W2T, GT = base + FS_GT;

  // Env-sharded ranks size the row lists for the worst case (learner.cu: 1.5 x the mean + 256 rows); the tiles
  // beyond this minibatch's actual row count have nothing to do except zeroing their partial sums.
  if (tile * 128 >= min(*p.count, p.cap)) {
    griddep_wait();                                     // the previous optimizer step may still be reading the partials
    float* part = p.part + static_cast<size_t>(tile) * p.part_stride;
    for (int i = threadIdx.x; i < H * G.aout; i += FS_THREADS) part[G.po_w2 + i] = 0.f;
    if (static_cast<int>(threadIdx.x) < G.aout) part[G.po_b2 + threadIdx.x] = 0.f;
    if (net == 0 && static_cast<int>(threadIdx.x) < G.aout) part[p.po_logstd + threadIdx.x] = 0.f;
    if (threadIdx.x == 0) part[G.po_loss] = 0.f;
    return;
  }
  if (threadIdx.x == FS_WORKERS) {
    FS_STAMP(16);
    mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1);
    mbar_init(&empty_bar[0], 1); mbar_init(&empty_bar[1], 1);
    mbar_init(accf0, 1); mbar_init(accf1, 1); mbar_init(headf, 1); mbar_init(bwdf, 1); mbar_init(dh1f, 1);
    mbar_init(xfull, FS_WORKERS);
    mbar_init(h2r, 8); mbar_init(gr, 8);
    for (int b = 0; b < 4; ++b) { mbar_init(&h1r[b], 8); mbar_init(&dz2r[b], 8); }
    mbar_init(&w0x[0], 1); mbar_init(&w0x[1], 1);
    fence_mbar_init();
  }
  if (warp == FS_MMA_WARP) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc0 = tmem_base, acc1 = tmem_base + 256;
  const uint32_t acc_head = acc1;               // [256, 288): head outputs, hi | lo halves (after acc1 is drained)
  const uint32_t acc_dw = acc1 + 32;            // [288, 288 + 32 * ceil(H/128)): head-kernel gradient, hi | lo halves

  if (warp == FS_TMA_WARP) {
    // ===================== weight producer: W0^T, W1^T, W1 k-blocks through the ring ==========
    if (elect_one()) {
      tma_prefetch_desc(&G.tm_w0); tma_prefetch_desc(&G.tm_w1); tma_prefetch_desc(&G.tm_w1k);
      griddep_wait();                 // the weight images are rewritten by the previous optimizer step
      const uint32_t bytes = static_cast<uint32_t>(H) * 128u;
      // Ring items: W0 k-blocks 0, 1, then W1 (L2), then W1 (dH1).  W0 k-blocks 2.. do not wait for a ring slot:
      // they are parked in R0, which nothing touches before epilogue 1, so the whole L1 GEMM is fed up front.
      const int nring0 = nk0 < 2 ? nk0 : 2;
      const int total = nring0 + 2 * nkH;
      for (int i = 0; i < total; ++i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        if (i < nring0 + nkH) {         // forward: k-block of W0 / W1 as stored, one {64 out, 64 in} box per 64 outputs
          const CUtensorMap* m = i < nring0 ? &G.tm_w0 : &G.tm_w1;
          const int kb = i < nring0 ? i : i - nring0;
          for (int c = 0; c < nkH; ++c) tma_load_2d(RB + s * FS_BSTAGE + c * 8192, m, &full_bar[s], c * 64, kb * 64);
        } else {                        // dH1: k-block (64 outputs) of W1 for all H inputs
          tma_load_2d(RB + s * FS_BSTAGE, &G.tm_w1k, &full_bar[s], (i - nring0 - nkH) * 64, 0);
        }
        if (i == nring0 - 1) {
          for (int kb = 2; kb < nk0; ++kb) {
            mbar_arrive_expect_tx(&w0x[kb - 2], bytes);
            for (int c = 0; c < nkH; ++c) tma_load_2d(R0 + (kb - 2) * FS_BSTAGE + c * 8192, &G.tm_w0, &w0x[kb - 2], c * 64, kb * 64);
          }
        }
      }
    }
  } else if (warp == FS_MMA_WARP) {
    // ===================== MMA issuer ==========================================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(H), 0u, 0u);
      int i = 0;
      // a_ready: per-k-block barriers of the A operand (the epilogue that produces A publishes it in
      // 64-column blocks, so this GEMM starts while the previous epilogue is still running)
      const uint32_t idesc_bmn = umma_idesc_bf16(128, static_cast<uint32_t>(H), 0u, 1u);     // B as stored: MN-major
      auto gemm = [&](uint32_t a_base, int nk, uint32_t acc, uint64_t* a_ready, bool b_mn) {
        for (int kb = 0; kb < nk; ++kb, ++i) {
          const int s = i & 1;
          const uint32_t ph = (i >> 1) & 1;
          if (a_ready) mbar_wait_spin(&a_ready[kb], 0);
          mbar_wait_spin(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = a_base + kb * 16384, sb = RB + s * FS_BSTAGE;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_bf16(acc, umma_smem_desc(sa + j * 32, 16, 1024),
                      b_mn ? umma_smem_desc(sb + j * 2048, 8192, 1024) : umma_smem_desc(sb + j * 32, 16, 1024),
                      b_mn ? idesc_bmn : idesc, (kb > 0 || j > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
      };
      FS_STAMP(17);
      mbar_wait_spin(xfull, 0);
      tc_fence_after();
      FS_STAMP(18);
      gemm(R1, nk0 < 2 ? nk0 : 2, acc0, nullptr, true);    // L1: X W0, k-blocks 0, 1 from the ring
      for (int kb = 2; kb < nk0; ++kb) {                   // ... k-blocks 2, 3 parked in R0
        mbar_wait_spin(&w0x[kb - 2], 0);
        tc_fence_after();
        const uint32_t sa = R1 + kb * 16384, sb = R0 + (kb - 2) * FS_BSTAGE;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          umma_bf16(acc0, umma_smem_desc(sa + j * 32, 16, 1024), umma_smem_desc(sb + j * 2048, 8192, 1024), idesc_bmn, 1u);
      }
      umma_commit(accf0);
      FS_STAMP(19);
      FS_STAMP(20);
      gemm(R0, nkH, acc1, h1r, true);        // L2: H1 W1
      umma_commit(accf1);
      FS_STAMP(21);
      // ---- head forward: [out_hi | out_lo][128 x 32] = H2 [W2_hi | W2_lo]; A = H2 K-major, B = W2^T K-major with the
      //      bf16 hi / lo halves stacked along N (the workers add the two 16-column halves)
      mbar_wait_spin(h2r, 0);
      tc_fence_after();
      {
        const uint32_t idesc_h = umma_idesc_bf16(128, 32u, 0u, 0u);
        uint32_t accum = 0;
        for (int kb = 0; kb < nkH; ++kb)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            umma_bf16(acc_head, umma_smem_desc(R1 + kb * 16384 + j * 32, 16, 1024),
                      umma_smem_desc(W2T + kb * 4096 + j * 32, 16, 1024), idesc_h, accum);
            accum = 1;
          }
        umma_commit(headf);
      }
      FS_STAMP(24);
      // ---- backward through the head --------------------------------------------------------------
      mbar_wait_spin(gr, 0);
      tc_fence_after();
      {
        // dA2[128 x H] = g W2^T, K = 16: A = g^T (MN-major, 64-row panels 4 KB apart), B = W2^T (MN-major, 64-column
        // panels 4 KB apart); the lo halves sit 16 rows = 2 KB into each panel
        const uint32_t idesc_a = umma_idesc_bf16(128, static_cast<uint32_t>(H), 1u, 1u);
        umma_bf16(acc0, umma_smem_desc(GT, 4096, 1024), umma_smem_desc(W2T, 4096, 1024), idesc_a, 0u);            // hi * hi
        umma_bf16(acc0, umma_smem_desc(GT, 4096, 1024), umma_smem_desc(W2T + 2048, 4096, 1024), idesc_a, 1u);     // hi * lo
        umma_bf16(acc0, umma_smem_desc(GT + 2048, 4096, 1024), umma_smem_desc(W2T, 4096, 1024), idesc_a, 1u);     // lo * hi
        // [dW2_hi | dW2_lo][c][j] = sum_r H2[r][c] g[r][j]: A = H2 (MN-major: M = c, K = rows), B = g^T (K-major, N = 32)
        const uint32_t idesc_w = umma_idesc_bf16(128, 32u, 1u, 0u);
        for (int mh = 0; mh < (H + 127) / 128; ++mh) {
#pragma unroll
          for (int t = 0; t < 8; ++t)
            umma_bf16(acc_dw + 32 * mh, umma_smem_desc(R1 + 2 * mh * 16384 + t * 2048, 16384, 1024),
                      umma_smem_desc(GT + (t >> 2) * 4096 + (t & 3) * 32, 16, 1024), idesc_w, t > 0 ? 1u : 0u);
        }
        umma_commit(bwdf);
      }
      FS_STAMP(25);
      FS_STAMP(22);
      // accumulates into acc1: every worker warp has read the head / dW2 columns of acc1 before its first
      // arrival on dz2r[0], while acc0 (dA2) is still being drained by the dZ2 epilogue
      gemm(R1, nkH, acc1, dz2r, false);      // dH1: dZ2 W1^T
      umma_commit(dh1f);
      FS_STAMP(23);
    }
  } else {
    // ===================== workers ==============================================================
    const int wt = static_cast<int>(threadIdx.x);               // 0..255
    const int ww = warp;                                         // 0..7
    const int q = warp & 3, hf = ww >> 2;
    const int erow = q * 32 + lane;                              // epilogue row == TMEM lane
    const int act = G.act, aout = G.aout;
    const int gchunk = wt & 7, grow0 = wt >> 3;                  // gather mapping: 8 lanes per 128-byte line, rows grow0 + 32 g
    if (wt == 0) FS_STAMP(0);

    // ---- gather the observation rows of this tile into R1 ---------------------------------------
    int src_g[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) src_g[g] = p.rowidx[tile * 128 + grow0 + 32 * g];
    const int lrow = tile * 128 + erow;                          // loss row of this thread (hf == 0 warps)
    const int count = min(*p.count, p.cap);
    const bool live = (hf == 0) && (lrow < count);
    const int src_l = live ? p.rowidx[lrow] : 0;
    // one warp instruction copies 4 rows x 128 contiguous bytes (4 L1 wavefronts; a lane-per-row mapping needed 32)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int row = grow0 + 32 * g;
      const __nv_bfloat16* src = p.obs_img + static_cast<size_t>(src_g[g]) * p.Dp + gchunk * 8;
      const uint32_t dst = R1 + row * 128 + ((gchunk ^ (row & 7)) << 4);
      for (int kb = 0; kb < nk0; ++kb) cp_async_16(dst + kb * 16384, src + kb * 64);
    }
    cp_async_commit();
    if (wt == 0) FS_STAMP(26);
    const float adv_sum = *p.adv_sum, adv_sq = *p.adv_sq;
    if (wt == 0) FS_STAMP(27);
    // ---- everything below reads what the previous optimizer step wrote (PDL: see common.cuh) -------
    griddep_wait();
    if (wt == 0) griddep_launch();
    if (wt == 0) FS_STAMP(28);
    // ---- small operands: biases, head bias / log_std, head kernel^T image (bf16 hi / lo, swizzled) --
    // (all global loads first: the st.shared wrappers are ordering barriers for the compiler)
    const float b0v = wt < H ? __ldcg(G.b0 + wt) : 0.f;             // H <= 256 == FS_WORKERS
    const float b1v = wt < H ? __ldcg(G.b1 + wt) : 0.f;
    float hbv = 0.f;
    if (wt < 16) hbv = wt < aout ? __ldcg(G.b2 + wt) : 0.f;
    else if (wt < 32) hbv = (net == 0 && wt - 16 < aout) ? __ldcg(p.log_std + wt - 16) : 0.f;
    for (int i = wt; i < 1024; i += FS_WORKERS) cp_async_16(W2T + i * 16, G.w2img + i);
    cp_async_commit();
    if (wt == 0) FS_STAMP(29);
    // the L1 GEMM needs only the gathered rows: publish them before the (slower) staging of the small operands
    cp_async_wait<1>();
    fence_proxy_async_smem();
    mbar_arrive(xfull);
    if (wt == 0) FS_STAMP(31);
    if (wt < H) { bias_s[wt] = b0v; bias_s[256 + wt] = b1v; }
    if (wt < 32) hb[wt] = hbv;
    if (wt >= 16 && wt < 32) hb[16 + wt] = 1.f / expf(hbv);          // 1 / scale (unused columns: 1)
    if (net == 0 && wt == 32) {
      // distrax: log|det| = sum log|scale|, scale = exp(log_std); summed in index order (train.py:223)
      float ls[FS_AP], logdet = 0.f;
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) ls[j] = j < aout ? __ldcg(p.log_std + j) : 0.f;
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) if (j < aout) logdet += logf(fabsf(expf(ls[j])));
      hb[48] = logdet;
    }
    if (wt == 0) FS_STAMP(30);
    cp_async_wait<0>();
    worker_bar();                                                // biases / hb / W2T visible to all workers; X complete
    if (wt == 0) FS_STAMP(1);
    if (net == 0 && ww == 0 && lane == 0) {
      for (int kb = 0; kb < nk0; ++kb) tma_store_2d(R1 + kb * 16384, &p.tm_xg, kb * 64, tile * 128);
      tma_store_commit();
    }
    // ---- epilogue 1: H1 = act(acc0 + b0) -> R0, then TMA store to HBM ------------------------------
    mbar_wait(accf0, 0);
    tc_fence_after();
    if (wt == 0) FS_STAMP(2);
    // the X store has read R1 before this warp's h1r arrivals let the L2 GEMM (and then epilogue 2, which
    // overwrites R1) proceed
    if (net == 0 && ww == 0 && lane == 0) tma_store_wait_read0();
    for (int b = 0; b < nkH; ++b) {
      epilogue_act(acc0, R0, bias_s, act, erow, q, b * 64 + hf * 32, 32);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h1r[b]);
    }
    if (wt == 0) FS_STAMP(3);
    if (ww == 0 && lane == 0) {
      for (int b = 0; b < nkH; ++b) {
        mbar_wait(&h1r[b], 0);
        tma_store_2d(R0 + b * 16384, &G.tm_h1, b * 64, tile * 128);
      }
      tma_store_commit();
    }

    // per-row loss inputs, requested here, in the shadow of the L2 GEMM's tail (32 distinct lines per load
    // instruction: ~2.5k cycles of load/store-unit time that must not sit in front of a GEMM or an epilogue)
    float in0 = 0.f, in1 = 0.f, actn[FS_AP];
#pragma unroll
    for (int j = 0; j < FS_AP; ++j) actn[j] = 0.f;
    if (live) {
      if (net == 0) {
        in0 = p.logp_old[src_l]; in1 = p.adv[src_l];
        const float* ap = p.action + static_cast<size_t>(src_l) * aout;
        if ((aout & 1) == 0) {                                     // rows are 8-byte aligned: half the load wavefronts
#pragma unroll
          for (int j = 0; j < FS_AP; j += 2)
            if (j < aout) { const float2 v = *reinterpret_cast<const float2*>(ap + j); actn[j] = v.x; actn[j + 1] = v.y; }
        } else {
#pragma unroll
          for (int j = 0; j < FS_AP; ++j) if (j < aout) actn[j] = ap[j];
        }
      } else {
        in0 = p.v_old[src_l]; in1 = p.tgt[src_l];
      }
    }


    // ---- epilogue 2: H2 = act(acc1 + b1) -> R1 -------------------------------------------------------
    mbar_wait(accf1, 0);
    tc_fence_after();
    if (wt == 0) FS_STAMP(4);
    epilogue_act(acc1, R1, bias_s + 256, act, erow, q, hf * (H >> 1), H >> 1);
    fence_proxy_async_smem();                                    // H2 (and W2T) -> async proxy for the head MMAs
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(h2r);
    if (wt == 0) FS_STAMP(5);

    // ---- loss and gradient seed g = dL/dout: thread = row (hf == 0 warps) ---------------------------
    mbar_wait(headf, 0);
    tc_fence_after();
    if (wt == 0) FS_STAMP(6);
    float dls[FS_AP], g[FS_AP];
    float s_loss = 0.f;
#pragma unroll
    for (int j = 0; j < FS_AP; ++j) { dls[j] = 0.f; g[j] = 0.f; }
    if (hf == 0) {
      float out[FS_AP];
      {
        float o2[32];
        tmem_ld_32x32(acc_head + (static_cast<uint32_t>(q * 32) << 16), o2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < FS_AP; ++j) out[j] = o2[j] + o2[FS_AP + j];      // W2_hi and W2_lo contributions
      }
      if (live) {
        const float inv_n = p.inv_mb;
        if (net == 0) {
          // distrax MultivariateNormalDiag: z = (a - loc) * (1/scale); train.py:223, 234-239
          float z[FS_AP], inv_s[FS_AP];
          float quad = 0.f;
#pragma unroll
          for (int j = 0; j < FS_AP; ++j) {
            z[j] = 0.f; inv_s[j] = 0.f;
            if (j < aout) {
              inv_s[j] = hb[32 + j];
              const float mean = out[j] + hb[j];
              z[j] = (actn[j] - mean) * inv_s[j];
              quad += -0.5f * z[j] * z[j] - 0.91893853320467274178f;
            }
          }
          const float logdet = hb[48];
          const float logp = quad - logdet;
          const float ratio = expf(logp - in0);
          const float adv_mean = adv_sum * inv_n;
          const float adv_std = sqrtf(adv_sq * inv_n);
          const float adv = (in1 - adv_mean) / (adv_std + 1e-8f);                 // train.py:235
          const float lo = 1.f - p.clip_eps, hi = 1.f + p.clip_eps;
          const float l1 = ratio * adv;
          const float l2 = fminf(fmaxf(ratio, lo), hi) * adv;
          s_loss = fminf(l1, l2);
          const float w1 = l1 < l2 ? 1.f : (l1 == l2 ? 0.5f : 0.f);
          const float dmin = (w1 + (1.f - w1) * dclip_f(ratio, lo, hi)) * adv;
          const float g_logp = -inv_n * dmin * ratio;
#pragma unroll
          for (int j = 0; j < FS_AP; ++j) {
            g[j] = g_logp * (z[j] * inv_s[j]);
            dls[j] = j < aout ? g_logp * (z[j] * z[j] - 1.f) : 0.f;
          }
        } else {
          // clipped value loss, train.py:226-231
          const float v = out[0] + hb[0];
          const float dvv = v - in0;
          const float v_clip = in0 + fminf(fmaxf(dvv, -p.clip_eps), p.clip_eps);
          const float e1 = v - in1, e2 = v_clip - in1;
          const float vl = e1 * e1, vlc = e2 * e2;
          s_loss = fmaxf(vl, vlc);
          const float wa = vl > vlc ? 1.f : (vl == vlc ? 0.5f : 0.f);
          g[0] = p.vf_coef * 0.5f * inv_n * (wa * 2.f * e1 + (1.f - wa) * 2.f * e2 * dclip_f(dvv, -p.clip_eps, p.clip_eps));
        }
      }
      // g^T as bf16 hi/lo, [16 j][128 rows] SW128 (the MN-major A of dA2 and the K-major B of dW2)
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) {
        uint32_t hi, lo;
        split_bf16(g[j], hi, lo);
        sts_u16(GT + sw32_off(j, erow), hi);
        sts_u16(GT + sw32_off(16 + j, erow), lo);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(gr);
    // tile sums (fixed order: lanes, then warps): dlog_std, head bias gradient, loss term
#pragma unroll
    for (int j = 0; j < FS_AP; ++j) { dls[j] = warp_sum(dls[j]); g[j] = warp_sum(g[j]); }
    s_loss = warp_sum(s_loss);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) { red[ww * 40 + j] = dls[j]; red[ww * 40 + 16 + j] = g[j]; }
      red[ww * 40 + 32] = s_loss;
    }
    worker_bar();
    if (wt == 0) FS_STAMP(7);
    float* part = p.part + static_cast<size_t>(tile) * p.part_stride;
    if (wt <= 32) {
      float s = 0.f;
      for (int w = 0; w < 4; ++w) s += red[w * 40 + wt];         // hf == 0 warps are ww 0..3
      if (wt == 32) part[G.po_loss] = s;
      else if (wt >= 16) { if (wt - 16 < aout) part[G.po_b2 + wt - 16] = s; }
      else if (net == 0 && wt < aout) part[p.po_logstd + wt] = s;
    }

    // ---- dW2 (head kernel gradient) and dZ2 = dA2 * f'(H2) ----------------------------------------
    mbar_wait(bwdf, 0);
    tc_fence_after();
    if (wt == 0) FS_STAMP(8);
    if (hf * 128 < H) {                                          // warp (q, hf) reads the c-tile mh = hf
      float dw[FS_AP];
      {
        float d2[32];
        tmem_ld_32x32(acc_dw + 32 * hf + (static_cast<uint32_t>(q * 32) << 16), d2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < FS_AP; ++j) dw[j] = d2[j] + d2[FS_AP + j];       // g_hi and g_lo contributions
      }
      const int c = hf * 128 + erow;
      if (c < H) {
#pragma unroll
        for (int j = 0; j < FS_AP; ++j) if (j < aout) part[G.po_w2 + c * aout + j] = dw[j];
      }
    }
    for (int b = 0; b < nkH; ++b) {
      epilogue_dact(acc0, R1, R1, act, erow, q, b * 64 + hf * 32, 32);            // in place: H2 -> dZ2
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dz2r[b]);
    }
    if (wt == 0) FS_STAMP(9);
    if (ww == 0 && lane == 0) {
      for (int b = 0; b < nkH; ++b) {
        mbar_wait(&dz2r[b], 0);
        tma_store_2d(R1 + b * 16384, &G.tm_dz2, b * 64, tile * 128);
      }
      tma_store_commit();
    }

    // ---- epilogue 3: dZ1 = acc1 * f'(H1) -> R1 -> TMA store ------------------------------------------
    mbar_wait(dh1f, 0);                                            // dH1 MMAs done: R1 (dZ2) no longer read by UMMA
    tc_fence_after();
    if (wt == 0) FS_STAMP(10);
    if (ww == 0 && lane == 0) tma_store_wait_read0();              // ... nor by the dZ2 TMA store
    worker_bar();
    for (int b = 0; b < nkH; ++b) {                                // 64-column blocks, each stored as soon as it is complete
      epilogue_dact(acc1, R0, R1, act, erow, q, b * 64 + hf * 32, 32);
      fence_proxy_async_smem();
      worker_bar();
      if (ww == 0 && lane == 0) tma_store_2d(R1 + b * 16384, &G.tm_dz1, b * 64, tile * 128);
    }
    if (wt == 0) FS_STAMP(11);
    if (ww == 0 && lane == 0) {
      tma_store_commit();
      tma_store_wait_read0();                                      // smem may be released once the bulk stores have read it;
    }                                                              // their global writes complete with the grid
    if (wt == 0) FS_STAMP(12);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FS_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace minppo
