// fp4.cuh -- NVFP4 (e2m1 codes, e4m3 per-16 scales) device helpers shared by the bf16 fast kernels (sm_100a).
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace b200q {
namespace fp4 {
using namespace fast;

__device__ __forceinline__ uint32_t cvt_e4m3x2(float hi, float lo) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_e2m1x2(float hi, float lo) {  // byte: lo element in bits 0-3
    uint16_t r;
    asm("{ .reg .b8 t; cvt.rn.satfinite.e2m1x2.f32 t, %1, %2; cvt.u16.u8 %0, t; }" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}

// eight fp32 values (four f32x2 pairs, element order) -> one word of eight e2m1 nibbles, element 0 in bits 0-3.  Written as one
// PTX block so the four byte results are merged by the conversion itself instead of being masked and permuted one by one.
__device__ __forceinline__ uint32_t cvt_e2m1x8(f32x2 p0, f32x2 p1, f32x2 p2, f32x2 p3) {
    uint32_t r;
    asm("{\n"
        ".reg .b8 t0, t1, t2, t3;\n"
        ".reg .f32 a0, a1, a2, a3, a4, a5, a6, a7;\n"
        "mov.b64 {a0, a1}, %1;\n"
        "mov.b64 {a2, a3}, %2;\n"
        "mov.b64 {a4, a5}, %3;\n"
        "mov.b64 {a6, a7}, %4;\n"
        "cvt.rn.satfinite.e2m1x2.f32 t0, a1, a0;\n"
        "cvt.rn.satfinite.e2m1x2.f32 t1, a3, a2;\n"
        "cvt.rn.satfinite.e2m1x2.f32 t2, a5, a4;\n"
        "cvt.rn.satfinite.e2m1x2.f32 t3, a7, a6;\n"
        "mov.b32 %0, {t0, t1, t2, t3};\n"
        "}\n"
        : "=r"(r) : "l"(p0.v), "l"(p1.v), "l"(p2.v), "l"(p3.v));
    return r;
}

// fast path: reciprocal normal (s_eff >= 2^-100) and no quotient of a non-zero bf16 (>= 2^-133) underflows to zero
// (s_eff <= 2^16): an underflowed -0.0 would keep its sign through the fused "+ 0.0", the reference's two-step
// (divide, then add the zero-point) turns it into +0.0.
__device__ __forceinline__ bool fp4_scale_is_safe(float s_eff) { return s_eff >= 7.8886090522101181e-31f && s_eff <= 65536.0f; }

static __device__ __noinline__ uint32_t fix_group_fp4(const uint4 raw, float s_eff, uint32_t packed) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    Bracket br;
    br.init(s_eff);
    float rl, rh, dummy;
    unpack2(br.lo, rl, dummy);
    unpack2(br.hi, rh, dummy);
    const bool all = !fp4_scale_is_safe(s_eff);
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        const uint32_t half = (e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16);
        const float x = __uint_as_float(half);
        const uint32_t ca = cvt_e2m1x2(0.0f, __fmaf_rn(x, rl, 0.0f)), cb = cvt_e2m1x2(0.0f, __fmaf_rn(x, rh, 0.0f));
        if (all || ca != cb) packed = (packed & ~(0xfu << (4 * e))) | (quant_fp4(x, s_eff) << (4 * e));
    }
    return packed;
}

struct Fp4Entry { float r_lo, r_hi, s_eff, unsafe; };


// exact bf16(absmax / 6) from the bf16 |max| bits (<< 16): bracketed constant reciprocal, IEEE fallback when the two ends
// round apart or the operand is outside [2^-100, 2^100] (zero is fine: both products are 0)
__device__ __forceinline__ float fp4_loc_scale(uint32_t abits) {
    const float a = __uint_as_float(abits);
    const float lo = __fmul_rn(a, (1.0f / 6.0f) * 0.99999952316284179688f), hi = __fmul_rn(a, (1.0f / 6.0f) * 1.00000047683715820312f);
    const uint32_t u = cvt_bf16x2(hi, lo);
    const bool in_range = (abits - 0x0d800000u) <= (0x71800000u - 0x0d800000u) || abits == 0;
    if (in_range && (u >> 16) == (u & 0xffffu)) return __uint_as_float(u << 16);
    return round_to<DT_BF16>(__fdiv_rn(a, 6.0f));
}

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

}  // namespace fp4
}  // namespace b200q
