// fastmath.cuh -- sm_100a packed-math building blocks of the issue-tuned bf16 kernels.
//
// Exact T(x / s) without an IEEE division per element ("bracketed reciprocal"):
//   r = rcp.approx(s) (<= 1 ulp), r_lo = r(1 - 2^-21), r_hi = r(1 + 2^-21)  =>  x*r_lo <= x/s <= x*r_hi in magnitude,
//   with >= 2^-22 relative slack on both sides after every rounding involved.  Rounding is monotone, so if the
//   two bracket ends round to the same low-precision value (bf16 / e2m1 code), the reference's
//   round(fp32(x/s)) rounds to it too.  When they differ (p ~ 1.5e-5 .. 2.5e-4 per element) the element is
//   recomputed with the IEEE chain from qmath.cuh.  Packed FMUL2/FFMA2 + F2FP keep this at 2.5 issue slots per
//   element instead of ~9 for div.rn.
#pragma once
#include "common.cuh"

namespace b200q {
namespace fast {

__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) { uint32_t r; asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { uint32_t r; asm("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmaxabs2(uint32_t a, uint32_t b) { uint32_t r; asm("max.xorsign.abs.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t cvt_bf16x2(float hi, float lo) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

struct f32x2 { uint64_t v; };
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
// a*b + 0.0: kills -0.0 exactly like the reference's "+ zero_point(0)" (x = -0.0 -> +0.0), same issue cost as mul2
__device__ __forceinline__ f32x2 mul2_plus0(f32x2 a, f32x2 b) {
    f32x2 r;
    const uint64_t z = 0;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(z));
    return r;
}
// the two bf16 halves of a 32-bit word as an fp32 pair (lo element first)
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t w) { return pack2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

// ---- ALU-pipe relief.  ncu on every bf16 fast kernel shows the half-rate ALU pipe (LOP3 / PRMT / F2FP / HMNMX2 / VIMNMX) at
// 60-70 % with the FMA pipe at 20-35 %: the kernels are bound by ALU issue, not by HBM.  These variants move work across:
// bf16 -> fp32 through the mixed-precision FMA (FHFMA.BF16 reads either half of the word directly: x * 1 + (-0.0) is exact for
// every x including -0.0), instead of IMAD.SHL + LOP3
__device__ __forceinline__ f32x2 bf16x2_to_f32x2_fma(uint32_t w) {
    float lo, hi;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; fma.rn.f32.bf16 %0, l, %3, %4; fma.rn.f32.bf16 %1, h, %3, %4; }"
        : "=f"(lo), "=f"(hi) : "r"(w), "h"((unsigned short)0x3f80), "f"(-0.0f));
    return pack2(lo, hi);
}
// "do the two bracket ends round to the same bf16?" accumulated on the FMA pipe: acc += (a - b)^2 per half (HFMA2.BF16); a zero
// accumulator means every pair was equal.  (a - b)^2 can only underflow to zero for |a| < ~1e-17, where both ends quantize to
// the same code anyway; inf - inf gives NaN, which reads as "different".
__device__ __forceinline__ uint32_t hdiff2_acc(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    asm("fma.rn.bf16x2 %0, %1, %1, %2;" : "=r"(acc) : "r"(d), "r"(acc));
    return acc;
}
__device__ __forceinline__ bool hdiff2_any(uint32_t acc) { return (acc & 0x7fff7fffu) != 0; }
// x + (x >> 4) in one instruction (mad.hi -> LEA.HI): merges neighbouring nibbles-in-bytes
__device__ __forceinline__ uint32_t fold_nibbles(uint32_t x) {
    uint32_t r;
    asm("mad.hi.u32 %0, %1, %2, %1;" : "=r"(r) : "r"(x), "r"(0x10000000u));
    return r;
}

struct Bracket {
    f32x2 lo, hi;  // (r_lo, r_lo), (r_hi, r_hi)
    __device__ __forceinline__ void init(float s) {
        const float r = rcp_approx(s);
        const float a = __fmul_rn(r, 0.99999952316284179688f);  // 1 - 2^-21
        const float b = __fmul_rn(r, 1.00000047683715820312f);  // 1 + 2^-21
        lo = pack2(a, a);
        hi = pack2(b, b);
    }
    // single evaluation (bf16 value / bf16 scale: see group_tma_kernel's ONE note): lo = hi = (r, r)
    __device__ __forceinline__ void init1(float s) {
        const float r = rcp_approx(s);
        lo = pack2(r, r);
        hi = lo;
    }
};
// exact T(a / divisor) for a non-negative bf16 statistic and a compile-time divisor, without an IEEE division on the
// common path: multiply by the bracketed constant reciprocal, fall back when the two ends round differently.
__device__ __forceinline__ float div_const_bf16(float a, float divisor) {
    const float c = 1.0f / divisor;  // folded at compile time (round-to-nearest)
    const float lo = __fmul_rn(a, c * 0.99999952316284179688f), hi = __fmul_rn(a, c * 1.00000047683715820312f);
    const uint32_t u = cvt_bf16x2(hi, lo);
    const bool ok = ((u >> 16) == (u & 0xffffu)) && (a == 0.0f || (a >= 7.8886090522101181e-31f && a <= 1.2676506002282294e30f));
    if (ok) return __uint_as_float(u << 16);
    return round_to<DT_BF16>(__fdiv_rn(a, divisor));
}

// the same for callers that have proven the quotient can never sit within 2^-21 of a bf16 rounding boundary: a has an 8-bit
// significand and the divisor at most 9 bits (7.5, 15, 448, 6), so a / divisor stays >= 2^-13 (relative) away from every boundary
__device__ __forceinline__ float div_const_bf16_1(float a, float divisor) {
    const float c = 1.0f / divisor;  // folded at compile time (round-to-nearest)
    if (a == 0.0f || (a >= 7.8886090522101181e-31f && a <= 1.2676506002282294e30f)) return round_to<DT_BF16>(__fmul_rn(a, c));
    return round_to<DT_BF16>(__fdiv_rn(a, divisor));
}

// 2^-100 <= s <= 1: reciprocal normal, quotients of in-group data neither overflow nor lose their sign
__device__ __forceinline__ bool scale_is_safe(uint32_t s_bits) { return (s_bits - 0x0d800000u) <= (0x3f800000u - 0x0d800000u); }

}  // namespace fast
}  // namespace b200q
