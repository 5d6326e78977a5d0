// minppo_b200 -- fused minibatch forward + PPO loss + backward-to-dZ for the 2-hidden-layer nets
// (the reference's default model.num_layers = 2, /root/reference/minppo/config.py:53).
//
// One CTA = one 128-row tile of the minibatch x one net (actor or critic).  Everything between
// the gathered observation rows and the layer gradients dZ stays on chip:
//
//   gather X (cp.async by row index, SW128)          -> smem R1
//   L1: acc0 = X  W0^T   (tcgen05, B streamed by TMA)  -> epilogue: +b0, act, bf16 -> smem R0 (H1) -> TMA store
//   L2: acc1 = H1 W1^T                                  -> epilogue: +b1, act, bf16 -> smem R1 (H2)
//   head (fp32 SIMT): out = H2 W2 + b2; Gaussian log-prob / clipped surrogate (actor CTA) or
//        clipped value loss (critic CTA) -> g = dL/dout  (train.py:218-243)
//   dZ2 = (g W2^T) * f'(H2) in place in R1 -> TMA store; dW2 / db2 / db1 / dlog_std / loss partials
//   dH1: acc0 = dZ2 W1     (tcgen05)                    -> epilogue: * f'(H1), bf16 -> smem R1 (dZ1) -> TMA store
//
// Warp roles: warp 0 TMA producer (weight k-blocks through a 2-stage ring), warp 1 TMEM
// allocator + MMA issuer, warps 2..9 workers (gather, epilogues, SIMT head), 2 warps per TMEM
// lane quadrant so that every scheduler has two epilogue warps in flight.
// H1, dZ2 and dZ1 are also written to HBM (bf16) because the split-K weight-gradient GEMM
// (umma_gemm.cuh, EPI_PARTIAL) runs as its own launch over all tiles.
#pragma once

#include "common.cuh"
#include "minppo_internal.h"
#include "umma_gemm.cuh"

namespace minppo {

constexpr int FS_THREADS = 320;
constexpr int FS_WORKERS = 256;
constexpr int FS_AP = 16;                               // padded head width (A <= 16)
constexpr int FS_R0 = 0;                                // H1            64 KB
constexpr int FS_R1 = 65536;                            // X / H2 / dZ2 / dZ1  64 KB
constexpr int FS_RB = 131072;                           // weight ring   2 x 32 KB
constexpr int FS_BSTAGE = 32768;
constexpr int FS_MISC = 196608;
constexpr int FS_W2S = FS_MISC;                         // [256][16] f32
constexpr int FS_BIAS = FS_W2S + 256 * FS_AP * 4;       // [2][256] f32
constexpr int FS_GS = FS_BIAS + 2 * 256 * 4;            // [128][16] f32
constexpr int FS_CS = FS_GS + 128 * FS_AP * 4;          // [4][256] f32
constexpr int FS_RED = FS_CS + 4 * 256 * 4;             // [8][24] f32
constexpr int FS_BARS = FS_RED + 8 * 24 * 4;            // mbarriers + tmem slot
constexpr int FS_SMEM_BYTES = FS_BARS + 128 + 1024;     // + alignment slack

struct alignas(64) FusedNet {
  CUtensorMap tm_w0t;            // W0^T image [H][Dp]   box {64, H}
  CUtensorMap tm_w1t;            // W1^T image [H][H]    box {64, H}
  CUtensorMap tm_w1n;            // W1   image [H][H]    box {64, H}   (n = in, k = out)
  CUtensorMap tm_h1;             // act[1] [M_pad][H]    box {64, 128}  (TMA store)
  CUtensorMap tm_dz2;            // dz[2]
  CUtensorMap tm_dz1;            // dz[1]
  const float* b0;               // arena pointers
  const float* b1;
  const float* w2;               // head kernel [H][aout]
  const float* b2;               // head bias [aout]
  float* colsum;                 // [m_tiles][H] column sums of dZ1 (bias gradient of layer 0)
  int act;                       // ACT_*
  int aout;                      // A (actor) or 1 (critic)
  int po_w2, po_b2, po_bh, po_loss;   // offsets of this net's fields inside a head partial
};

struct alignas(64) FusedParams {
  FusedNet net[2];
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
};

MINPPO_DEVINL void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

MINPPO_DEVINL uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
MINPPO_DEVINL void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)) : "memory");
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
MINPPO_DEVINL float act_deriv(float h, int act) { return act == ACT_RELU ? (h > 0.f ? 1.f : 0.f) : (1.f - h * h); }
MINPPO_DEVINL float dclip_f(float x, float lo, float hi) {
  return (x > lo && x < hi) ? 1.f : ((x == lo || x == hi) ? 0.5f : 0.f);
}

// accumulator (32 columns per chunk) -> +bias, activation, bf16 -> swizzled smem tile
template <int ACT>
MINPPO_DEVINL void epilogue_act_t(uint32_t tmem_acc, uint32_t dst_base, const float* bias_s, int row, int q, int col0,
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

__global__ void __launch_bounds__(FS_THREADS, 1) fused_step_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* w2s = reinterpret_cast<float*>(sm + FS_W2S);
  float* bias_s = reinterpret_cast<float*>(sm + FS_BIAS);       // [0..256) layer 0, [256..512) layer 1
  float* gs = reinterpret_cast<float*>(sm + FS_GS);
  float* cs = reinterpret_cast<float*>(sm + FS_CS);
  float* red = reinterpret_cast<float*>(sm + FS_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + FS_BARS);
  uint64_t* full_bar = bars;            // [2]
  uint64_t* empty_bar = bars + 2;       // [2]
  uint64_t* accf = bars + 4;            // [2]
  uint64_t* xfull = bars + 6;
  uint64_t* h1r = bars + 7;
  uint64_t* dz2r = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int net = static_cast<int>(blockIdx.x) / p.m_tiles;
  const int tile = static_cast<int>(blockIdx.x) % p.m_tiles;
  const FusedNet& G = p.net[net];
  const int H = p.H, nkH = H >> 6, nk0 = p.Dp >> 6;
  const uint32_t R0 = base + FS_R0, R1 = base + FS_R1, RB = base + FS_RB;

  if (threadIdx.x == 0) {
    mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1);
    mbar_init(&empty_bar[0], 1); mbar_init(&empty_bar[1], 1);
    mbar_init(&accf[0], 1); mbar_init(&accf[1], 1);
    mbar_init(xfull, FS_WORKERS);
    mbar_init(h1r, 8);
    mbar_init(dz2r, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc0 = tmem_base, acc1 = tmem_base + 256;

  if (warp == 0) {
    // ===================== weight producer: W0^T, W1^T, W1 k-blocks through the ring ==========
    if (elect_one()) {
      tma_prefetch_desc(&G.tm_w0t); tma_prefetch_desc(&G.tm_w1t); tma_prefetch_desc(&G.tm_w1n);
      const uint32_t bytes = static_cast<uint32_t>(H) * 128u;
      const int total = nk0 + 2 * nkH;
      for (int i = 0; i < total; ++i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        const CUtensorMap* m = i < nk0 ? &G.tm_w0t : (i < nk0 + nkH ? &G.tm_w1t : &G.tm_w1n);
        const int kb = i < nk0 ? i : (i < nk0 + nkH ? i - nk0 : i - nk0 - nkH);
        tma_load_2d(RB + s * FS_BSTAGE, m, &full_bar[s], kb * 64, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ==========================================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(H), 0u, 0u);
      int i = 0;
      auto gemm = [&](uint32_t a_base, int nk, uint32_t acc) {
        for (int kb = 0; kb < nk; ++kb, ++i) {
          const int s = i & 1;
          const uint32_t ph = (i >> 1) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = a_base + kb * 16384, sb = RB + s * FS_BSTAGE;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_bf16(acc, umma_smem_desc(sa + j * 32, 16, 1024), umma_smem_desc(sb + j * 32, 16, 1024), idesc,
                      (kb > 0 || j > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
      };
      mbar_wait(xfull, 0);
      tc_fence_after();
      gemm(R1, nk0, acc0);            // L1: X W0^T
      umma_commit(&accf[0]);
      mbar_wait(h1r, 0);
      tc_fence_after();
      gemm(R0, nkH, acc1);            // L2: H1 W1^T
      umma_commit(&accf[1]);
      mbar_wait(dz2r, 0);
      tc_fence_after();
      gemm(R1, nkH, acc0);            // dH1: dZ2 W1
      umma_commit(&accf[0]);
    }
  } else {
    // ===================== workers ==============================================================
    const int wt = static_cast<int>(threadIdx.x) - 64;          // 0..255
    const int ww = warp - 2;                                     // 0..7
    const int q = warp & 3, hf = ww >> 2;
    const int erow = q * 32 + lane;                              // epilogue row == TMEM lane
    const int act = G.act, aout = G.aout;
    const int AQ = (aout + 3) >> 2;
    const int srow = wt >> 1, shalf = wt & 1;                    // SIMT (row, half) mapping
    const int grow = tile * 128 + srow;

    // ---- gather the observation rows of this tile into R1 ---------------------------------------
    const int src = p.rowidx[grow];
    for (int kb = shalf; kb < nk0; kb += 2)
      gather_line(R1 + kb * 16384, srow, p.obs_img + static_cast<size_t>(src) * p.Dp + kb * 64);
    cp_async_commit();
    // small operands
    for (int i = wt; i < H * FS_AP; i += FS_WORKERS) {
      const int k = i >> 4, j = i & 15;
      w2s[i] = j < aout ? G.w2[k * aout + j] : 0.f;
    }
    for (int i = wt; i < H; i += FS_WORKERS) { bias_s[i] = G.b0[i]; bias_s[256 + i] = G.b1[i]; }
    // per-row loss inputs, prefetched (used after the second epilogue)
    const int count = min(*p.count, p.cap);
    const bool live = (shalf == 0) && (grow < count);
    float in0 = 0.f, in1 = 0.f, actn[FS_AP];
#pragma unroll
    for (int j = 0; j < FS_AP; ++j) actn[j] = 0.f;
    if (live) {
      if (net == 0) {
        in0 = p.logp_old[src]; in1 = p.adv[src];
#pragma unroll
        for (int j = 0; j < FS_AP; ++j) if (j < aout) actn[j] = p.action[static_cast<size_t>(src) * aout + j];
      } else {
        in0 = p.v_old[src]; in1 = p.tgt[src];
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    mbar_arrive(xfull);
    worker_bar();                                                // w2s / bias visible to all workers

    // ---- epilogue 1: H1 = act(acc0 + b0) -> R0, then TMA store to HBM ------------------------------
    mbar_wait(&accf[0], 0);
    tc_fence_after();
    epilogue_act(acc0, R0, bias_s, act, erow, q, hf * (H >> 1), H >> 1);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(h1r);
    if (ww == 0 && lane == 0) {
      mbar_wait(h1r, 0);
      for (int s = 0; s < nkH; ++s) tma_store_2d(R0 + s * 16384, &G.tm_h1, s * 64, tile * 128);
      tma_store_commit();
    }

    // ---- epilogue 2: H2 = act(acc1 + b1) -> R1 -------------------------------------------------------
    mbar_wait(&accf[1], 0);
    tc_fence_after();
    epilogue_act(acc1, R1, bias_s + 256, act, erow, q, hf * (H >> 1), H >> 1);
    worker_bar();

    // ---- head: out[j] = sum_c H2[row][c] w2[c][j]   (thread = (row, half of the columns)) ----------
    float acc[FS_AP];
#pragma unroll
    for (int j = 0; j < FS_AP; ++j) acc[j] = 0.f;
    {
      const int nch = H >> 4;                                    // 16-byte chunks per half row
      const int cbeg = shalf * (H >> 1);
      const int rot = (shalf && (((cbeg >> 3) & 4) == 0)) ? 4 : 0;   // keep the two halves on different banks
      for (int t = 0; t < nch; ++t) {
        int tt = t + rot;
        if (tt >= nch) tt -= nch;
        const int c = cbeg + 8 * tt;
        const uint4 hv = lds128(R1 + sw_off(srow, c));
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float h = (e & 1) ? bf16_hi(hw[e >> 1]) : bf16_lo(hw[e >> 1]);
          const float4* wr = reinterpret_cast<const float4*>(w2s + (c + e) * FS_AP);
#pragma unroll
          for (int jq = 0; jq < FS_AP / 4; ++jq) {
            if (jq < AQ) {
              const float4 wv = wr[jq];
              acc[4 * jq] = fmaf(h, wv.x, acc[4 * jq]);
              acc[4 * jq + 1] = fmaf(h, wv.y, acc[4 * jq + 1]);
              acc[4 * jq + 2] = fmaf(h, wv.z, acc[4 * jq + 2]);
              acc[4 * jq + 3] = fmaf(h, wv.w, acc[4 * jq + 3]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
    }

    // ---- loss and gradient seed g = dL/dout, one thread per row (half 0) -------------------------
    float dls[FS_AP];
    float s_loss = 0.f;
#pragma unroll
    for (int j = 0; j < FS_AP; ++j) dls[j] = 0.f;
    if (shalf == 0) {
      float g[FS_AP];
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) g[j] = 0.f;
      if (live) {
        const float inv_n = p.inv_mb;
        if (net == 0) {
          // distrax MultivariateNormalDiag: z = (a - loc) * (1/scale); train.py:223, 234-239
          float z[FS_AP], inv_s[FS_AP];
          float quad = 0.f, logdet = 0.f;
#pragma unroll
          for (int j = 0; j < FS_AP; ++j) {
            z[j] = 0.f; inv_s[j] = 0.f;
            if (j < aout) {
              const float scale = expf(p.log_std[j]);
              inv_s[j] = 1.f / scale;
              const float mean = acc[j] + G.b2[j];
              z[j] = (actn[j] - mean) * inv_s[j];
              quad += -0.5f * z[j] * z[j] - 0.91893853320467274178f;
              logdet += logf(fabsf(scale));
            }
          }
          const float logp = quad - logdet;
          const float ratio = expf(logp - in0);
          const float adv_mean = (*p.adv_sum) * inv_n;
          const float adv_std = sqrtf((*p.adv_sq) * inv_n);
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
          const float v = acc[0] + G.b2[0];
          const float dvv = v - in0;
          const float v_clip = in0 + fminf(fmaxf(dvv, -p.clip_eps), p.clip_eps);
          const float e1 = v - in1, e2 = v_clip - in1;
          const float vl = e1 * e1, vlc = e2 * e2;
          s_loss = fmaxf(vl, vlc);
          const float wa = vl > vlc ? 1.f : (vl == vlc ? 0.5f : 0.f);
          g[0] = p.vf_coef * 0.5f * inv_n * (wa * 2.f * e1 + (1.f - wa) * 2.f * e2 * dclip_f(dvv, -p.clip_eps, p.clip_eps));
        }
      }
#pragma unroll
      for (int jq = 0; jq < FS_AP / 4; ++jq)
        *reinterpret_cast<float4*>(gs + srow * FS_AP + 4 * jq) = make_float4(g[4 * jq], g[4 * jq + 1], g[4 * jq + 2], g[4 * jq + 3]);
    }
    // tile sums of dlog_std and of the loss term (fixed order: lanes, then warps)
#pragma unroll
    for (int j = 0; j < FS_AP; ++j) dls[j] = warp_sum(dls[j]);
    s_loss = warp_sum(s_loss);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) red[ww * 24 + j] = dls[j];
      red[ww * 24 + FS_AP] = s_loss;
    }
    worker_bar();

    float* part = p.part + static_cast<size_t>(tile) * p.part_stride;
    if (wt <= FS_AP) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w * 24 + wt];
      if (wt == FS_AP) part[G.po_loss] = s;
      else if (net == 0 && wt < aout) part[p.po_logstd + wt] = s;
    }
    if (wt >= 32 && wt < 32 + aout) {                              // head bias gradient: sum_r g[r][j]
      const int j = wt - 32;
      float s = 0.f;
      for (int r = 0; r < 128; ++r) s += gs[r * FS_AP + j];
      part[G.po_b2 + j] = s;
    }

    // ---- thread = hidden column c: dZ2 in place, dW2 (head kernel grad), db1 ---------------------
    for (int c = wt; c < H; c += FS_WORKERS) {
      float w[FS_AP], dw[FS_AP];
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) { w[j] = w2s[c * FS_AP + j]; dw[j] = 0.f; }
      float db = 0.f;
      const uint32_t cbase = R1 + static_cast<uint32_t>((c >> 6) * 16384 + (c & 7) * 2);
      const int pos = (c & 63) >> 3;
#pragma unroll 4
      for (int r = 0; r < 128; ++r) {
        const uint32_t addr = cbase + r * 128 + ((pos ^ (r & 7)) << 4);
        const float h = __uint_as_float(lds_u16(addr) << 16);
        float da = 0.f;
#pragma unroll
        for (int jq = 0; jq < FS_AP / 4; ++jq) {
          if (jq < AQ) {
            const float4 gv = *reinterpret_cast<const float4*>(gs + r * FS_AP + 4 * jq);
            da = fmaf(gv.x, w[4 * jq], da); da = fmaf(gv.y, w[4 * jq + 1], da);
            da = fmaf(gv.z, w[4 * jq + 2], da); da = fmaf(gv.w, w[4 * jq + 3], da);
            dw[4 * jq] = fmaf(h, gv.x, dw[4 * jq]); dw[4 * jq + 1] = fmaf(h, gv.y, dw[4 * jq + 1]);
            dw[4 * jq + 2] = fmaf(h, gv.z, dw[4 * jq + 2]); dw[4 * jq + 3] = fmaf(h, gv.w, dw[4 * jq + 3]);
          }
        }
        const __nv_bfloat16 dz = __float2bfloat16_rn(da * act_deriv(h, act));
        sts_u16(addr, static_cast<uint32_t>(__bfloat16_as_ushort(dz)));
        db += __bfloat162float(dz);
      }
#pragma unroll
      for (int j = 0; j < FS_AP; ++j) if (j < aout) part[G.po_w2 + c * aout + j] = dw[j];
      part[G.po_bh + c] = db;
    }
    fence_proxy_async_smem();
    tc_fence_before();                                             // acc0 reads of epilogue 1 ordered before the dH1 MMAs
    __syncwarp();
    if (lane == 0) mbar_arrive(dz2r);
    if (ww == 0 && lane == 0) {
      mbar_wait(dz2r, 0);
      for (int s = 0; s < nkH; ++s) tma_store_2d(R1 + s * 16384, &G.tm_dz2, s * 64, tile * 128);
      tma_store_commit();
    }

    // ---- epilogue 3: dZ1 = acc0 * f'(H1) -> R1 -> TMA store; column sums -> db0 --------------------
    mbar_wait(&accf[0], 1);                                        // dH1 MMAs done: R1 (dZ2) no longer read by UMMA
    tc_fence_after();
    if (ww == 0 && lane == 0) tma_store_wait_read0();              // ... nor by the dZ2 TMA store
    worker_bar();
    {
      const uint32_t taddr = acc0 + (static_cast<uint32_t>(q * 32) << 16);
      const int col0 = hf * (H >> 1);
      for (int c0 = col0; c0 < col0 + (H >> 1); c0 += 32) {
        float v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = sw_off(erow, c0 + 8 * j);
          const uint4 hh = lds128(R0 + off);
          const uint32_t hw[4] = {hh.x, hh.y, hh.z, hh.w};
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int e = 8 * j + 2 * t;
            v[e] *= act_deriv(bf16_lo(hw[t]), act);
            v[e + 1] *= act_deriv(bf16_hi(hw[t]), act);
            w[t] = pack_bf16x2(v[e], v[e + 1]);
            v[e] = bf16_lo(w[t]); v[e + 1] = bf16_hi(w[t]);        // bias gradient sums the stored (rounded) dZ
          }
          sts128(R1 + off, make_uint4(w[0], w[1], w[2], w[3]));
        }
        const float csum = warp_colsum32(v);
        cs[q * 256 + c0 + lane] = csum;
      }
    }
    fence_proxy_async_smem();
    worker_bar();
    if (ww == 0 && lane == 0) {
      for (int s = 0; s < nkH; ++s) tma_store_2d(R1 + s * 16384, &G.tm_dz1, s * 64, tile * 128);
      tma_store_commit();
    }
    for (int c = wt; c < H; c += FS_WORKERS)
      G.colsum[static_cast<size_t>(tile) * H + c] = (cs[c] + cs[256 + c]) + (cs[512 + c] + cs[768 + c]);
    if (ww == 0 && lane == 0) tma_store_wait_all0();               // all bulk stores complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace minppo
