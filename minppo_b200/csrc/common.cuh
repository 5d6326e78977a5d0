// minppo_b200 -- common device helpers for sm_100a (B200).
// Thin inline-PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM),
// cp.async, plus small math/bf16 utilities.  No CUTLASS dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace minppo {

#define MINPPO_DEVINL __device__ __forceinline__

// ------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------
MINPPO_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

MINPPO_DEVINL uint32_t lane_id() { return threadIdx.x & 31u; }

MINPPO_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

MINPPO_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------
// programmatic dependent launch (PDL).  Protocol used by the per-minibatch kernels:
//   prologue + reads of update-static data;  griddep_wait();  griddep_launch();  main body.
// Triggering only AFTER the own wait keeps completion transitive: when kernel K+1 starts, K has
// passed its wait, hence K-1 has completed.  Both are no-ops for a launch without the attribute.
// Data another kernel of the chain rewrites (params, partials) is read with ld.global.cg after
// the wait: with PDL the L1 invalidate of the launch boundary may precede those writes.
// ------------------------------------------------------------------------------------
MINPPO_DEVINL void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
MINPPO_DEVINL void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------
MINPPO_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
MINPPO_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
MINPPO_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MINPPO_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// suspend-time hint of mbarrier.try_wait (ns): without it the attempt may return after an implementation-defined, short time and
// the waiting warp re-issues it in a loop on the scheduler it shares with working warps
#ifndef MINPPO_TRY_WAIT_HINT_NS
#define MINPPO_TRY_WAIT_HINT_NS 1000000u
#endif
MINPPO_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(MINPPO_TRY_WAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug (missing arrive / wrong tx count) traps with an error after
// ~2 s instead of hanging the GPU.  try_wait suspends the warp in hardware for a bounded time per attempt; the clock is
// only read every 64 attempts, so that a waiting warp costs its scheduler next to nothing (16 worker warps wait most of
// the time, on the schedulers the working warps issue from).
#ifndef MINPPO_WAIT_SLEEP_NS
#define MINPPO_WAIT_SLEEP_NS 0
#endif
template <bool POLITE = true>
MINPPO_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      if (mbar_try_wait(bar, parity)) return;
#if MINPPO_WAIT_SLEEP_NS
      if (POLITE) asm volatile("nanosleep.u32 %0;" ::"r"(static_cast<uint32_t>(MINPPO_WAIT_SLEEP_NS)));
#endif
    }
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// Wait of the single MMA-issuing thread.  MINPPO_MMA_SPIN = 1: busy-polling mbarrier.test_wait (never suspends: round 1
// measured ~2k idle cycles in front of a GEMM after a suspended try_wait, with the clock read in every iteration) -- but a
// busy loop takes issue slots from the four epilogue warps that share its scheduler; 0 (default): the suspending try_wait
// loop above.  Measured in round 2 (same box, configs[1]): 5.880 ms (spin) vs 5.830 ms (suspend) per update.
#ifndef MINPPO_MMA_SPIN
#define MINPPO_MMA_SPIN 0
#endif
MINPPO_DEVINL void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
#if MINPPO_MMA_SPIN
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
      uint32_t ok;
      asm volatile(
          "{\n\t.reg .pred P;\n\t"
          "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, P;\n\t}\n"
          : "=r"(ok)
          : "r"(addr), "r"(parity)
          : "memory");
      if (ok) return;
    }
    if (clock64() - t0 > 4000000000LL) __trap();
  }
#else
  mbar_wait<false>(bar, parity);
#endif
}

// generic-proxy writes (st.shared / cp.async) -> visible to the async proxy (UMMA / TMA)
MINPPO_DEVINL void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------
// cp.async (LDGSTS), 16-byte
// ------------------------------------------------------------------------------------
MINPPO_DEVINL void cp_async_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
MINPPO_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
MINPPO_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------
// TMA: 2-D tiled load, global -> shared, completion on an mbarrier
// ------------------------------------------------------------------------------------
MINPPO_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
MINPPO_DEVINL void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA stores: shared -> global, bulk-group completion
MINPPO_DEVINL void tma_store_2d(uint32_t smem_src, const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
MINPPO_DEVINL void tma_store_3d(uint32_t smem_src, const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 1-D bulk store shared -> global (16-byte aligned addresses, size a multiple of 16), part of the thread's bulk async-group
MINPPO_DEVINL void bulk_store_1d(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
MINPPO_DEVINL void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
MINPPO_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
MINPPO_DEVINL void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
MINPPO_DEVINL void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// async-proxy (TMA) global writes -> ordered before later generic-proxy operations of this thread
// (release to other CTAs of the same grid through a grid barrier)
MINPPO_DEVINL void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

MINPPO_DEVINL void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
MINPPO_DEVINL uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// ------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, loads, fences
// ------------------------------------------------------------------------------------
MINPPO_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
MINPPO_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MINPPO_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
MINPPO_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MINPPO_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16 in, fp32 accumulate); one thread issues.
MINPPO_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
MINPPO_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns; thread i gets lane (base+i).
MINPPO_DEVINL void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
MINPPO_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp, PTX ISA "tcgen05 matrix descriptor")
// ------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, version 1 (sm_100).
//   K-major  tile [rows][64 bf16]: 8-row groups 1024 B apart -> SBO = 1024; LBO unused (=1).
//   MN-major tile [k rows][64 bf16] per 64-wide MN chunk: 8-k groups 1024 B apart -> SBO = 1024,
//             next MN chunk `lbo_bytes` away.
MINPPO_DEVINL uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);              // [0,14)  start address
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;         // [16,30) leading byte offset
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;         // [32,46) stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                                  // [46,48) version = 1
  d |= static_cast<uint64_t>(2) << 61;                                  // [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: BF16 x BF16 -> FP32, dense.
MINPPO_DEVINL constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                 uint32_t b_mn_major) {
  return (1u << 4)                 // c_format = F32
         | (1u << 7)               // a_format = BF16
         | (1u << 10)              // b_format = BF16
         | (a_mn_major << 15)      // a_major
         | (b_mn_major << 16)      // b_major
         | ((N >> 3) << 17)        // n_dim
         | ((M >> 4) << 24);       // m_dim
}

// ------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------
MINPPO_DEVINL float fast_tanh(float x) {            // MUFU.TANH, |rel err| ~ 2^-11
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh through ex2/rcp (2 MUFU): abs err ~1e-7, well below one bf16 ulp of the stored output
MINPPO_DEVINL float exp_tanh(float x) {
  // 1 - 2 / (1 + exp(2x)): ex2.approx saturates to +inf / 0, rcp.approx(+inf) = 0, so the
  // limits +-1 come out without clamping.  5 instructions, 2 of them MUFU.
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

MINPPO_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// x = hi + lo with hi, lo bf16 (relative error 2^-17)
MINPPO_DEVINL void split_bf16(float x, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}
// byte offset of element (row rr, column c) inside a [2 ap][ncols] bf16 SW128 tile set: one panel of 2 ap rows x 128 B per
// 64 columns.  The head operands (width padded to ap = 16 or 32) are stored as a bf16 hi / lo pair stacked along the rows:
// rr = part * ap + j.
MINPPO_DEVINL uint32_t swp_off(int ap, int rr, int c) {
  return static_cast<uint32_t>((c >> 6) * (ap * 256) + rr * 128 + ((((c & 63) >> 3) ^ (rr & 7)) << 4) + (c & 7) * 2);
}
MINPPO_DEVINL float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
MINPPO_DEVINL float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

}  // namespace minppo
