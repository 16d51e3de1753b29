// hostmath.cu -- host-only instantiation of qmath.cuh (the arithmetic the kernels inline), exported through a
// small C ABI so the CPU test-suite can check it against the oracle without a GPU.  Not used by the product.
#include <stdint.h>
#include <string.h>
#include "qmath.cuh"

using namespace b200q;

namespace {
template <int DT> float ld(const void* p, int64_t i) {
    if (DT == DT_BF16) { uint32_t u = ((uint32_t)((const uint16_t*)p)[i]) << 16; float f; memcpy(&f, &u, 4); return f; }
    if (DT == DT_F16) return __half2float(((const __half*)p)[i]);
    return ((const float*)p)[i];
}
template <int DT> void st(void* p, int64_t i, float v) {
    if (DT == DT_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
    else if (DT == DT_F16) ((__half*)p)[i] = __float2half_rn(v);
    else ((float*)p)[i] = v;
}

// group-wise fused compress on the host, element order identical to the kernels' per-chunk math
template <int DT>
int compress_group(const void* w, int64_t rows, int64_t cols, int qt, int nbits, int sym, int g, const float* gs_in,
                   uint8_t* codes /*one code per element*/, void* scale /*T or e4m3 bytes*/, int8_t* zp, float* gs_out) {
    const float lo = qt == QT_INT ? -(float)(1 << (nbits - 1)) : (qt == QT_FP8 ? -448.0f : -6.0f);
    const float hi = qt == QT_INT ? (float)((1 << (nbits - 1)) - 1) : (qt == QT_FP8 ? 448.0f : 6.0f);
    float gs = 1.0f;
    if (qt == QT_FP4) {
        if (gs_in) gs = gs_in[0];
        else {
            float mn = INFINITY, mx = -INFINITY;
            for (int64_t i = 0; i < rows * cols; i++) { float v = ld<DT>(w, i); mn = fminf(mn, v); mx = fmaxf(mx, v); }
            gs = gparam<DT>(fmaxf(fabsf(fminf(mn, 0.0f)), fabsf(fmaxf(mx, 0.0f))));
        }
        if (gs_out) gs_out[0] = gs;
    }
    const int64_t G = cols / g;
    for (int64_t r = 0; r < rows; r++)
        for (int64_t k = 0; k < G; k++) {
            float mn = INFINITY, mx = -INFINITY, a = 0.0f;
            for (int64_t c = k * g; c < (k + 1) * g; c++) {
                float v = ld<DT>(w, r * cols + c);
                mn = fminf(mn, v); mx = fmaxf(mx, v); a = fmaxf(a, fabsf(v));
            }
            float s = 1.0f, z = 0.0f, s_eff = 1.0f;
            if (qt == QT_INT && !sym) qparams_asym<DT>(mn, mx, lo, hi, s, z);
            else if (qt == QT_FP4) ((uint8_t*)scale)[r * G + k] = qparams_fp4<DT>(a, gs, s_eff);
            else s = scale_sym<DT>(a, qt == QT_INT ? (hi - lo) * 0.5f : hi);
            if (qt != QT_FP4) st<DT>(scale, r * G + k, s);
            if (zp) zp[r * G + k] = (int8_t)(int)z;
            for (int64_t c = k * g; c < (k + 1) * g; c++) {
                float v = ld<DT>(w, r * cols + c);
                uint8_t code;
                if (qt == QT_INT) code = (uint8_t)(int8_t)quant_int<DT>(v, s, z, !sym, lo, hi);
                else if (qt == QT_FP8) code = quant_fp8<DT>(v, s, true);
                else code = (uint8_t)quant_fp4(v, s_eff);
                codes[r * cols + c] = code;
            }
        }
    return 0;
}
template <int DT>
int fq_group(const void* w, int64_t rows, int64_t cols, int qt, int nbits, int g, const void* scale, const int8_t* zp,
             int has_zp, const float* gsp, void* out) {
    const float lo = qt == QT_INT ? -(float)(1 << (nbits - 1)) : (qt == QT_FP8 ? -448.0f : -6.0f);
    const float hi = qt == QT_INT ? (float)((1 << (nbits - 1)) - 1) : (qt == QT_FP8 ? 448.0f : 6.0f);
    const int64_t G = cols / g;
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) {
            const int64_t k = r * G + c / g;
            float s = ld<DT>(scale, k), z = zp ? (float)zp[k] : 0.0f, x = ld<DT>(w, r * cols + c), y;
            if (qt == QT_INT) y = fq_int<DT>(x, s, z, zp != nullptr, lo, hi);
            else if (qt == QT_FP8) y = fq_fp8<DT>(x, s, has_zp != 0);
            else y = fq_fp4<DT>(x, fdiv(s, gsp[0]));
            st<DT>(out, r * cols + c, y);
        }
    return 0;
}
}  // namespace

extern "C" {
int hm_compress_group(const void* w, int dt, int64_t rows, int64_t cols, int qt, int nbits, int sym, int g, const float* gs_in,
                      uint8_t* codes, void* scale, int8_t* zp, float* gs_out) {
    switch (dt) {
    case DT_BF16: return compress_group<DT_BF16>(w, rows, cols, qt, nbits, sym, g, gs_in, codes, scale, zp, gs_out);
    case DT_F16: return compress_group<DT_F16>(w, rows, cols, qt, nbits, sym, g, gs_in, codes, scale, zp, gs_out);
    case DT_F32: return compress_group<DT_F32>(w, rows, cols, qt, nbits, sym, g, gs_in, codes, scale, zp, gs_out);
    }
    return -1;
}
int hm_fq_group(const void* w, int dt, int64_t rows, int64_t cols, int qt, int nbits, int g, const void* scale, const int8_t* zp,
                int has_zp, const float* gs, void* out) {
    switch (dt) {
    case DT_BF16: return fq_group<DT_BF16>(w, rows, cols, qt, nbits, g, scale, zp, has_zp, gs, out);
    case DT_F16: return fq_group<DT_F16>(w, rows, cols, qt, nbits, g, scale, zp, has_zp, gs, out);
    case DT_F32: return fq_group<DT_F32>(w, rows, cols, qt, nbits, g, scale, zp, has_zp, gs, out);
    }
    return -1;
}
float hm_gparam(float absmax, int dt) {
    return dt == DT_BF16 ? gparam<DT_BF16>(absmax) : (dt == DT_F16 ? gparam<DT_F16>(absmax) : gparam<DT_F32>(absmax));
}
uint8_t hm_e4m3_encode(float x) { return e4m3_encode(x); }
float hm_e4m3_decode(uint8_t c) { return e4m3_decode(c); }
}
