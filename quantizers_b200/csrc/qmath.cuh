// qmath.cuh -- per-element / per-group arithmetic of the quantization hot path, written once as
// __host__ __device__ so the exact same code is (a) inlined into the sm_100a kernels and (b) compiled for the
// host into libb200q_hostmath.so, where the CPU test-suite checks it against the oracle without a GPU.
//
// Semantics = compressed-tensors 0.15.0.1 CPU eager (one rounding to the tensor dtype T per ATen op):
//   calculate_qparams   CT:quantization/utils/helpers.py:50-137
//   generate_gparam     CT:quantization/utils/helpers.py:309-338
//   _quantize/_quantize_dequantize/_dequantize   CT:quantization/lifecycle/forward_helpers.py:176-268
//   round_to_quantized_type_args / cast_to_fp4    CT:quantization/quant_args.py:46-66,439-475
// All values of dtype T are carried in registers as fp32 (every bf16/fp16 value is exact in fp32).
// No fast-math: divisions are IEEE (__fdiv_rn), no FMA contraction across the reference's rounding points.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_fp4.h>
#include <stdint.h>
#include <math.h>

#define HD __host__ __device__ __forceinline__

namespace b200q {

enum : int { DT_BF16 = 0, DT_F16 = 1, DT_F32 = 2 };
enum : int { QT_INT = 0, QT_FP8 = 1, QT_FP4 = 2 };
enum : int { ST_TENSOR = 0, ST_CHANNEL = 1, ST_GROUP = 2, ST_BLOCK = 3 };

HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);  // never contracted into an FMA
#else
    return a + b;
#endif
}
HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
HD float frint(float a) {  // round half to even, like torch.round
#ifdef __CUDA_ARCH__
    return rintf(a);
#else
    return nearbyintf(a);
#endif
}

// ---- round an fp32 value to T (RNE) and back: the "one ATen op" rounding
template <int DT> HD float round_to(float x);
template <> HD float round_to<DT_BF16>(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
template <> HD float round_to<DT_F16>(float x) { return __half2float(__float2half_rn(x)); }
template <> HD float round_to<DT_F32>(float x) { return x; }

template <int DT> HD float eps_of() {  // torch.finfo(T).eps
    return DT == DT_BF16 ? 0.0078125f : (DT == DT_F16 ? 0.0009765625f : 1.1920928955078125e-07f);
}
template <int DT> HD float tiny_of() {  // torch.finfo(T).tiny
    return DT == DT_F16 ? 6.103515625e-05f : 1.1754943508222875e-38f;
}

// ---- e4m3fn
HD uint8_t e4m3_encode(float x) {  // RNE, input already clamped to +-448 by the caller (NaN -> 0x7f)
    return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3);
}
HD float e4m3_decode(uint8_t c) {
    uint32_t e = (c >> 3) & 0xfu, m = c & 7u;
    float v;
    if (e == 0) v = (float)m * 0.001953125f;
    else v = u2f(((e + 120u) << 23) | (m << 20));  // 2^(e-7) * (1 + m/8); 0x7f (NaN) never produced upstream
    return (c & 0x80u) ? -v : v;
}

// ---- e2m1 magnitude index, CT:quantization/quant_args.py:53-66 thresholds (== RNE on the e2m1 grid)
HD uint32_t e2m1_index(float a) {
    return (a > 0.25f) + (a >= 0.75f) + (a > 1.25f) + (a >= 1.75f) + (a > 2.5f) + (a >= 3.5f) + (a > 5.0f);
}
HD float e2m1_value(uint32_t idx) {
    // {0, .5, 1, 1.5, 2, 3, 4, 6}
    return idx < 2 ? 0.5f * (float)idx : u2f((((idx >> 1) + 126u) << 23) | ((idx & 1u) << 22));
}

// =====================================================================================  qparams
// bit_range/2 for symmetric: int4 7.5, int8 127.5, fp8 448, fp4 6 (helpers.py:76-86)
template <int DT> HD float scale_sym(float absmax, float half_range) {
    float s = round_to<DT>(fdiv(absmax, half_range));
    return s == 0.0f ? eps_of<DT>() : s;
}

// asymmetric INT (helpers.py:96-98,129-131).  Returns the eps-fixed scale; zp as an integer-valued float.
template <int DT> HD void qparams_asym(float mn, float mx, float bit_min, float bit_max, float& s_out, float& z_out) {
    mn = fminf(mn, 0.0f);
    mx = fmaxf(mx, 0.0f);
    float d = round_to<DT>(fadd(mx, -mn));
    float s = round_to<DT>(fdiv(d, bit_max - bit_min));
    float t = round_to<DT>(fdiv(mn, s));  // un-eps'ed scale: 0/0 = NaN for an all-zero group
    float z = round_to<DT>(fadd(bit_min, -t));
    if (z == z) z = fminf(fmaxf(z, bit_min), bit_max);
    z = (z == z) ? frint(z) : 0.0f;  // NaN -> int8 cast gives 0 on the reference's CPU path
    s_out = s == 0.0f ? eps_of<DT>() : s;
    z_out = fadd(z, 0.0f);  // the zero-point is stored as int8: a rounded -0.0 becomes +0 (matters for fake-quant signs of zero)
}

// NVFP4 local scale (helpers.py:86,101-126): loc = T(absmax/6); s = e4m3(clamp(gs * loc)); 0 -> 0.125
// Returns the e4m3 code; s_eff = fp32(s) / gs is the divisor _quantize uses (forward_helpers.py:226-229).
template <int DT> HD uint8_t qparams_fp4(float absmax, float gs, float& s_eff) {
    float loc = round_to<DT>(fdiv(absmax, 6.0f));
    float sf = fmul(gs, loc);
    sf = fminf(fmaxf(sf, -448.0f), 448.0f);
    uint8_t code = e4m3_encode(sf);
    float s = e4m3_decode(code);
    if (s == 0.0f) { s = 0.125f; code = 0x20; }
    s_eff = fdiv(s, gs);
    return code;
}

// generate_gparam: T(T(1/A) * 2688), A clamped to tiny, NaN/Inf -> 1 (helpers.py:325-338; Tensor.__rdiv__)
template <int DT> HD float gparam(float absmax) {
    float a = absmax < tiny_of<DT>() ? tiny_of<DT>() : absmax;  // torch.clamp(min=tiny) keeps NaN
    float r = round_to<DT>(fdiv(1.0f, a));
    float g = round_to<DT>(fmul(r, 2688.0f));
    if (!(fabsf(g) <= 3.4028234663852886e38f)) g = 1.0f;
    return g;
}

// =====================================================================================  quantize
// INT: T(x/s) -> T(+zp) -> clamp -> RNE.  Returns the signed code.
template <int DT> HD float quant_int_f(float x, float s, float z, bool add_zp, float lo, float hi) {
    float u = round_to<DT>(fdiv(x, s));
    if (add_zp) u = round_to<DT>(fadd(u, z));
    if (u == u) u = fminf(fmaxf(u, lo), hi);
    return frint(u);  // keeps -0.0 (torch.round(-0.3) == -0.0), which fake_quantize propagates
}
template <int DT> HD int quant_int(float x, float s, float z, bool add_zp, float lo, float hi) {
    float u = quant_int_f<DT>(x, s, z, add_zp, lo, hi);
    return (u == u) ? (int)u : 0;
}
// FP8: T(x/s) (+0 when a zero-point tensor is present) -> clamp +-448 -> e4m3 RNE (double rounding via T)
template <int DT> HD uint8_t quant_fp8(float x, float s, bool add_zp) {
    float u = round_to<DT>(fdiv(x, s));
    if (add_zp) u = fadd(u, 0.0f);
    if (u == u) u = fminf(fmaxf(u, -448.0f), 448.0f);
    return e4m3_encode(u);
}
// FP4 (NVFP4): fp32 x / s_eff (+0) -> clamp +-6 -> e2m1 RNE; sign bit iff the pre-round value is < 0.
HD uint32_t quant_fp4(float x, float s_eff) {
    float u = fadd(fdiv(x, s_eff), 0.0f);
    u = fminf(fmaxf(u, -6.0f), 6.0f);
    return e2m1_index(fabsf(u)) | (u < 0.0f ? 8u : 0u);
}

// =====================================================================================  fake-quant / dequant leg
// dequant = T_s(q) ; (- zp) ; * s  -- each rounded to the scale dtype ST (forward_helpers.py:203-210,258-266)
template <int ST> HD float dequant_val(float q, float s, float z, bool sub_zp) {
    float d = round_to<ST>(q);
    if (sub_zp) d = round_to<ST>(fadd(d, -z));
    return round_to<ST>(fmul(d, s));
}
template <int DT> HD float fq_int(float x, float s, float z, bool has_zp, float lo, float hi) {
    float q = quant_int_f<DT>(x, s, z, has_zp, lo, hi);
    return dequant_val<DT>(q, s, z, has_zp);
}
template <int DT> HD float fq_fp8(float x, float s, bool has_zp) {
    float q = e4m3_decode(quant_fp8<DT>(x, s, has_zp));
    return dequant_val<DT>(q, s, 0.0f, has_zp);
}
template <int DT> HD float fq_fp4(float x, float s_eff) {  // fp32 arithmetic, final cast to T
    uint32_t n = quant_fp4(x, s_eff);
    float q = e2m1_value(n & 7u);
    // torch: x * sign -> -0.0 for negatives rounding to 0; minus zp(0) gives +0; times s: sign irrelevant after T cast of 0
    if (n & 8u) q = -q;
    float d = fadd(q, -0.0f);
    return round_to<DT>(fmul(d, s_eff));
}

}  // namespace b200q
