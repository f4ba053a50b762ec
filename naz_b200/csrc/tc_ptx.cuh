// Thin inline-PTX wrappers for the sm_100a features the tcgen05 engine uses:
// mbarrier, cp.async.bulk (TMA bulk copy), tcgen05.{alloc,mma,commit,ld,fence}, proxy fences.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tcx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------- mbarrier ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// The spin loop lives INSIDE the asm block (as CUTLASS does): to the compiler the wait is straight-line
// code, so the surrounding region stays provably convergent and warp-uniform values (MMA descriptors,
// slot counters) can be kept in uniform registers.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------- TMA bulk copy (global -> shared, completes on an mbarrier) ----------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------- proxy / tcgen05 fences ----------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// ---------------- TMEM ----------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_out)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {   // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor, no swizzle ("interleave"), K-major: 8x8 core matrices of 128
// contiguous bytes; LBO = byte stride between core matrices along K, SBO = along M/N
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor; version field = 1 on sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor for kind::f16, A = B = fp16, D = fp32, both K-major, M = 128.
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same MMA, executed convergently by the whole warp with the instruction itself predicated on `elected`
// (so descriptors can stay in uniform registers instead of being R2UR-broadcast from a divergent branch).
__device__ __forceinline__ void mma_f16_ss_elect(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                                 uint32_t accumulate, uint32_t elected) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar, uint32_t elected) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(elected)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand read from tensor memory (lane = row, one 32-bit column = two consecutive-K fp16)
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM, 16 lanes x 4 columns twice: threads 0-15 write lanes base..base+15 at columns [col, col+4), threads 16-31
// the same lanes at columns [col+IMM, col+IMM+4)
template <int IMM>
__device__ __forceinline__ void tmem_st16x2_4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.16x32bx2.x4.b32 [%0], %1, {%2,%3,%4,%5};\n" ::"r"(taddr), "n"(IMM), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}\n"
      : "+r"(pred));
  return pred;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// 16 lanes x 8 columns, twice: threads 0-15 get lanes base..base+15 at columns [col, col+8), threads 16-31 the
// same lanes at columns [col+IMM, col+IMM+8)  (probed on B200: tools/tc_probe_ld16.cu).
template <int IMM>
__device__ __forceinline__ void tmem_ld16x2_8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr), "n"(IMM));
}
template <int IMM>
__device__ __forceinline__ void tmem_ld16x2_2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x2.b32 {%0,%1}, [%2], %3;\n" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr), "n"(IMM));
}
// Instruction descriptor for kind::f16, A = B = fp16, D = fp32, both K-major, M = 64.
__device__ __forceinline__ uint32_t make_idesc_f16_m64(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((64u >> 4) << 24);
}

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// ---------------- math ----------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
// tanh(x) = 1 - 2 / (2^(2 x log2 e) + 1); two MUFU ops, abs error ~1e-7 (ex2 / rcp are ~1-2 ulp).
__device__ __forceinline__ float tanh_fast(float x) {
  float e = ex2_approx(x * 2.885390081777927f);
  return fmaf(-2.f, rcp_approx(e + 1.f), 1.f);
}
// fp32 -> (hi, lo) fp16 pair with saturation (|x| > 65504 clamps instead of producing inf * 0 = NaN).
__device__ __forceinline__ uint32_t pack_hi2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(b), "f"(a));   // low half = a
  return r;
}
__device__ __forceinline__ void unpack2(uint32_t h2, float& a, float& b) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}\n"
      : "=f"(a), "=f"(b)
      : "r"(h2));
}

// ---------------- packed fp32 pairs (Blackwell FFMA2 / FADD2) ----------------
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;\n" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 8 pre-activations ALREADY scaled by 2 log2(e), as 4 packed pairs -> tanh -> fp16 hi / lo chunks (16 B each).
// tanh(x) = 1 - 2 / (2^s + 1), s = 2 x log2 e.  MUFU is the scarce pipe of this engine (16 lanes/clk/SM), so the
// four reciprocals of a group share ONE rcp: with y = 2^s + 1, r = 1 / (y0 y1 y2 y3), 1/y0 = r (y1 y3) y2, ...
// (three extra roundings, ~2 ulp on a value <= 1).  s is clamped to <= 30, where tanh already rounds to 1 in fp32,
// so the product of four y stays below 2^121.  1.25 MUFU + ~7 other issue slots per element.
__device__ __forceinline__ void tanh8_scaled(const uint64_t* s2, uint4& hi4, uint4& lo4) {
  uint32_t h[4], l[4];
  const uint64_t one2 = pk2(1.f, 1.f), m2 = pk2(-2.f, -2.f);
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float a, b, c, d;
    upk2(s2[2 * g], a, b);
    upk2(s2[2 * g + 1], c, d);
    const uint64_t y01 = add2(pk2(ex2_approx(fminf(a, 30.f)), ex2_approx(fminf(b, 30.f))), one2);
    const uint64_t y23 = add2(pk2(ex2_approx(fminf(c, 30.f)), ex2_approx(fminf(d, 30.f))), one2);
    float pac, pbd;
    upk2(mul2(y01, y23), pac, pbd);                       // (y0 y2, y1 y3)
    const float r = rcp_approx(pac * pbd);
    const uint64_t R = mul2(pk2(r, r), pk2(pbd, pac));    // (r y1 y3, r y0 y2)
    const uint64_t t01 = fma2(mul2(R, y23), m2, one2);    // 1 - 2 / y0, 1 - 2 / y1
    const uint64_t t23 = fma2(mul2(R, y01), m2, one2);    // 1 - 2 / y2, 1 - 2 / y3
    float fa, fb;
    upk2(t01, a, b);
    h[2 * g] = pack_hi2(a, b);
    unpack2(h[2 * g], fa, fb);
    upk2(sub2(t01, pk2(fa, fb)), a, b);
    l[2 * g] = pack_hi2(a, b);
    upk2(t23, a, b);
    h[2 * g + 1] = pack_hi2(a, b);
    unpack2(h[2 * g + 1], fa, fb);
    upk2(sub2(t23, pk2(fa, fb)), a, b);
    l[2 * g + 1] = pack_hi2(a, b);
  }
  hi4 = make_uint4(h[0], h[1], h[2], h[3]);
  lo4 = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace tcx
