// sm_100a primitives used by the fused tensor-core kernels: mbarrier, 1-D TMA bulk copies,
// tcgen05 (TMEM allocation, UMMA descriptors, MMA issue, commit, TMEM loads).  Inline PTX only.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // back off between polls: a spinning warp otherwise steals issue slots from the warps it is waiting for
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}

// wait with a long hardware suspend hint: the thread sleeps inside try_wait until the phase completes instead of
// burning issue slots in a poll / nanosleep loop (the other pipelines of the CTA need them)
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
  } while (ok == 0);
}

// latency-critical single-thread wait (the MMA-issuing lane): poll without backing off
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// MUFU.EX2 / MUFU.LG2 without the denormal fix-up code the non-ftz intrinsics carry
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// generic-proxy writes to shared memory -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- 1-D TMA bulk copy global -> shared, completion on an mbarrier -----------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- named barriers (sub-CTA sync) ----------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------------
// executed by ONE full warp; writes the TMEM base address to *smem_out
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ---- UMMA descriptors ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE canonical layouts (16-byte units; core matrix = 8 x 16 B,
// stored as 128 contiguous bytes):
//   K-major  operand [rows, K]: byte(r, k) = (r%8)*16 + (k*es)%16 + (r/8)*SBO + ((k*es)/16)*LBO
//   MN-major operand [K, rows]: byte(k, r) = (r*es)%16 + (k%8)*16 + ((r*es)/16)*SBO + (k/8)*LBO
// (so the SAME bytes are a K-major [rows=a, K=b] operand and an MN-major [K=a, rows=b] operand with
//  LBO and SBO exchanged - used to get every transpose of the backward pass for free).
// bits: [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout type = 0.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Instruction descriptor for kind::f16 (A/B = bf16 or f16, D = fp32), dense, no negate.
//   fmt: 0 = f16, 1 = bf16;  a_mn / b_mn: 1 = MN-major operand.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int fmt, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mbarrier arrives once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}


// Shared-memory loads through an explicit 32-bit shared-window address held in a register: keeps ptxas from
// re-deriving the window base (S2UR SR_CgaCtaId + address arithmetic) in front of every access of a hot loop.
// Read-only data only (not volatile: the compiler may hoist / reorder these loads).
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}

// ---- packed half-precision ShiftedSoftplus for the filter-MLP epilogue ---------------------------------------
// ssp(x) * c for two columns at once, entirely in f16x2 (one MUFU.EX2 and ten packed ALU operations per PAIR instead
// of two MUFU and ~10 ALU per element):  ssp(x) = max(x, 0) + ln(1 + t) - ln 2,  t = 2^(-|x| log2 e) in (0, 1],
// ln(1 + t) = t * P4(t) (degree-4 fit on [0, 1], 8e-5).  The result is kept in f16 (10 mantissa bits): measured rms error
// 3e-4 against 1.4e-3 for rounding the exact value to bf16, so the f16 a' operand is the more accurate one.
__device__ __forceinline__ uint32_t ssp_cutoff_f16x2(float x0, float x1, __half2 c) {
  const __half2 x = __floats2half2_rn(x0, x1);
  // h2exp2 (fp32 MUFU.EX2 with cuda_fp16's fix-up, then rounded to f16), NOT the native ex2.approx.f16x2: the native
  // one is ~20 % faster in this epilogue but doubles the error of the whole model (measured: embedding 5e-4 -> 1.2e-3,
  // gradients 5e-3 -> 1.1e-2 against the oracle)
  const __half2 t = h2exp2(__hmul2(__habs2(x), __float2half2_rn(-1.4426950408889634f)));
  __half2 p = __float2half2_rn(0.04106098f);
  p = __hfma2(p, t, __float2half2_rn(-0.15602058f));
  p = __hfma2(p, t, __float2half2_rn(0.30466648f));
  p = __hfma2(p, t, __float2half2_rn(-0.4963666f));
  p = __hfma2(p, t, __float2half2_rn(0.99988779f));
  const __half2 q = __hfma2(t, p, __float2half2_rn(-0.6931471805599453f));
  const __half2 s = __hadd2(__hmax2(x, __float2half2_rn(0.0f)), q);
  const __half2 r = __hmul2(s, c);
  return *reinterpret_cast<const uint32_t*>(&r);
}

}  // namespace tc
}  // namespace cmp
