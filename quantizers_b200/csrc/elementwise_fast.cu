// elementwise_fast.cu -- bf16 quantize / fake_quantize with CALLER-SUPPLIED qparams at HBM speed, any strategy whose qparam is
// constant over an aligned run of 8 elements (TENSOR, CHANNEL, GROUP with g % 8 == 0, BLOCK with bw % 8 == 0):
//   CT quantize       CT:quantization/lifecycle/forward.py:37-73   (INT4 -> int8 codes, FP8 -> e4m3)
//   CT fake_quantize  CT:quantization/lifecycle/forward.py:149-181 (quantize, then (q - zp) * scale in the scale dtype)
// This is what the registered compressors (naive / float-quantized `compress`) and forward_quantize call when an observer has
// already written weight_scale / weight_zero_point.  The generic kernel (quant_elementwise.cu) re-derives the qparam index with
// 64-bit divisions per ELEMENT, loads 2 bytes at a time and divides with the IEEE sequence: 0.2-0.4 of the HBM roofline.
//
// Here a warp walks a row in batches of U 16-byte chunks per lane, every load of a batch (weights, scale, zero point) issued before
// the first conversion; the quotient T(x / s) comes from the bracketed reciprocal of fastmath.cuh (both bracket ends round to the
// same bf16, otherwise -- or when the scale is outside the bracket's safe range, or the zero point outside [-8, 7] -- the chunk is
// redone with the exact qmath.cuh chain), rounding / clamping / offsetting are packed bf16x2 / s16x2 instructions as in
// quant_group_tma.cu, and the dequantize leg of fake_quantize is the decode_fast.cu arithmetic.
#include <cstdlib>
#include "common.cuh"
#include "fastmath.cuh"
#include "fp4.cuh"
#include "kernels.cuh"

namespace b200q {
namespace {
using namespace fast;
using fp4::cvt_e4m3x2;

constexpr int EW_PACK = 3;  // INT4 only: quantize + pack_to_int32 (offset-binary nibbles, 8 per word)

struct EwFast {
    const uint4* x;
    const uint16_t* scale;
    const int8_t* zp;
    void* out;
    int64_t rows;
    int cpr;               // 16-byte chunks per row
    int sh;                // chunk -> qparam column: c >> sh (31: always column 0)
    int64_t rows_per_q;    // rows sharing one qparam row (1, block height, or `rows` for TENSOR)
    int64_t q_row_stride;  // qparams per qparam row
};

__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void e4m3x2_to_f32(uint32_t two_bytes, float& lo, float& hi) {
    asm("{ .reg .b16 t, l, h; .reg .b32 r; cvt.u16.u32 t, %2; cvt.rn.f16x2.e4m3x2 r, t; mov.b32 {l, h}, r; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }"
        : "=f"(lo), "=f"(hi) : "r"(two_bytes));
}

// the exact chain for one chunk (rare): QUANT -> 8 code bytes in o.x, o.y; FQ -> 8 bf16 values in o
template <int QT, int OP>
__device__ __noinline__ uint4 exact_chunk(const uint4 raw, float s, float z, bool use_zp) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        const float x = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
        if (OP == EW_PACK) {
            o[0] |= ((uint32_t)(quant_int<DT_BF16>(x, s, z, use_zp, -8.0f, 7.0f) + 8) & 0xfu) << (4 * e);
        } else if (OP == EW_QUANT) {
            uint32_t c;
            if (QT == QT_INT) c = (uint32_t)quant_int<DT_BF16>(x, s, z, use_zp, -8.0f, 7.0f) & 0xffu;
            else c = quant_fp8<DT_BF16>(x, s, use_zp);
            o[e >> 2] |= c << (8 * (e & 3));
        } else {
            float y;
            if (QT == QT_INT) y = fq_int<DT_BF16>(x, s, z, use_zp, -8.0f, 7.0f);
            else y = fq_fp8<DT_BF16>(x, s, use_zp);
            o[e >> 1] |= (__float_as_uint(y) >> 16) << (16 * (e & 1));
        }
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// ZP: INT -> a zero-point tensor is supplied (added before rounding, subtracted when dequantizing); FP8 -> "+ 0" of the fp8 zero point
template <int QT, int OP, bool ZP, int U>
__global__ void __launch_bounds__(256) elementwise_fast_kernel(const EwFast p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const uint32_t kMagic = 0x43484348u, kUnbias = 0xbcc0bcc0u;  // bf16x2 (200, 200) / s16x2 (-0x4340): RNE, clamp [-8, 7], + 8
    for (int64_t row = warp; row < p.rows; row += nwarps) {
        const uint4* xrow = p.x + row * p.cpr;
        const int64_t qrow = (row / p.rows_per_q) * p.q_row_stride;
        for (int base = 0; base < p.cpr; base += 32 * U) {
            uint4 v[U];
            uint32_t sc[U];
            int zq[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int c = base + 32 * u + lane;
                const bool ok = c < p.cpr;
                const int64_t qi = qrow + (c >> p.sh);
                v[u] = ok ? ldg_stream(xrow + c) : make_uint4(0u, 0u, 0u, 0u);
                sc[u] = ok ? (uint32_t)__ldg(p.scale + qi) : 0x3f80u;
                zq[u] = (QT == QT_INT && ZP && ok) ? (int)__ldg(p.zp + qi) : 0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int c = base + 32 * u + lane;
                if (c >= p.cpr) continue;
                const float s = __uint_as_float(sc[u] << 16), z = (float)zq[u];
                Bracket br;
                br.init(s);
                const bool unsafe = !scale_is_safe(sc[u] << 16) || (QT == QT_INT && ZP && (zq[u] < -8 || zq[u] > 7));
                const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                const uint32_t z2 = (__float_as_uint(z) >> 16) * 0x10001u;
                const uint32_t c2 = (0x4308u + (uint32_t)zq[u]) * 0x10001u;  // bf16x2 (136 + zp): exact for zp in [-8, 7]
                const uint32_t s2 = sc[u] * 0x10001u;
                const bool zero_fix = !ZP || zq[u] == 0;  // a negative value that rounds to zero dequantizes to -0.0 (torch.round keeps the sign)
                uint32_t h[4], y[4], diff = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const f32x2 x = bf16x2_to_f32x2_fma(w[k]);
                    float al, ah, bl, bh;
                    if (QT == QT_INT) {
                        unpack2(mul2(x, br.lo), al, ah);
                        unpack2(mul2(x, br.hi), bl, bh);
                        uint32_t q = cvt_bf16x2(ah, al);
                        diff |= q ^ cvt_bf16x2(bh, bl);
                        if (ZP) q = hadd2(q, z2);
                        h[k] = __viaddmin_s16x2_relu(hadd2(q, kMagic), kUnbias, 0x000f000fu);  // code + 8 in [0, 15]
                        if (OP == EW_FQ) {
                            uint32_t d = hmul2(hsub2(0x43004300u | h[k], c2), s2);  // ((128 + code + 8) - (136 + zp)) * s
                            if (zero_fix) d |= q & ~((d & 0x7fff7fffu) + 0x7fff7fffu) & 0x80008000u;
                            y[k] = d;
                        }
                    } else {
                        unpack2(ZP ? mul2_plus0(x, br.lo) : mul2(x, br.lo), al, ah);
                        unpack2(ZP ? mul2_plus0(x, br.hi) : mul2(x, br.hi), bl, bh);
                        const uint32_t q = cvt_bf16x2(ah, al);
                        diff |= q ^ cvt_bf16x2(bh, bl);
                        float ql, qh;
                        unpack2(bf16x2_to_f32x2_fma(q), ql, qh);
                        h[k] = cvt_e4m3x2(qh, ql);
                        if (OP == EW_FQ) {
                            float dl, dh;
                            e4m3x2_to_f32(h[k], dl, dh);
                            y[k] = cvt_bf16x2(__fmul_rn(dh, s), __fmul_rn(dl, s));
                        }
                    }
                }
                uint4 o;
                if (OP == EW_FQ) o = make_uint4(y[0], y[1], y[2], y[3]);
                else if (OP == EW_PACK) {
                    const uint32_t x01 = prmt(h[0], h[1], 0x6420), x23 = prmt(h[2], h[3], 0x6420);  // code + 8 is the stored nibble
                    o = make_uint4(prmt(fold_nibbles(x01), fold_nibbles(x23), 0x6420), 0u, 0u, 0u);
                } else if (QT == QT_INT) {
                    // code + 8 -> two's-complement nibble (xor 8) -> sign-extended byte (bit 3 * 0x1e fills the high nibble)
                    const uint32_t n01 = prmt(h[0], h[1], 0x6420) ^ 0x08080808u, n23 = prmt(h[2], h[3], 0x6420) ^ 0x08080808u;
                    o = make_uint4(n01 | ((n01 & 0x08080808u) * 0x1eu), n23 | ((n23 & 0x08080808u) * 0x1eu), 0u, 0u);
                } else o = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), 0u, 0u);
                if (diff != 0 || unsafe) o = exact_chunk<QT, OP>(v[u], s, z, ZP);
                if (OP == EW_PACK) stg_stream((uint32_t*)p.out + row * p.cpr + c, o.x);
                else if (OP == EW_FQ) stg_stream((uint4*)p.out + row * p.cpr + c, o);
                else stg_stream((uint2*)p.out + row * p.cpr + c, make_uint2(o.x, o.y));
            }
        }
    }
}

// CT dequantize of un-packed INT codes (int8 -> bf16; naive / int-quantized `decompress`, CT:quantization/lifecycle/forward.py:77-145):
// (T(q) - T(zp)) * scale with one bf16 rounding per step.  A byte becomes its integer through the 2^23 mantissa trick
// (0x4B000000 | (q + 128) is the float 2^23 + q + 128), exact in bf16 for every int8.
template <bool ZP, int U>
__global__ void __launch_bounds__(256) dequant_int8_fast_kernel(const EwFast p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const uint2* codes = (const uint2*)p.x;
    for (int64_t row = warp; row < p.rows; row += nwarps) {
        const uint2* crow = codes + row * p.cpr;
        const int64_t qrow = (row / p.rows_per_q) * p.q_row_stride;
        for (int base = 0; base < p.cpr; base += 32 * U) {
            uint2 v[U];
            uint32_t sc[U];
            int zq[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int c = base + 32 * u + lane;
                const bool ok = c < p.cpr;
                const int64_t qi = qrow + (c >> p.sh);
                v[u] = ok ? __ldg(crow + c) : make_uint2(0u, 0u);
                sc[u] = ok ? (uint32_t)__ldg(p.scale + qi) : 0u;
                zq[u] = (ZP && ok) ? (int)__ldg(p.zp + qi) : 0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int c = base + 32 * u + lane;
                if (c >= p.cpr) continue;
                const uint32_t s2 = sc[u] * 0x10001u;
                const uint32_t z2 = (__float_as_uint((float)zq[u]) >> 16) * 0x10001u;  // every int8 is a bf16
                const uint32_t w[2] = {v[u].x ^ 0x80808080u, v[u].y ^ 0x80808080u};      // q + 128 per byte
                uint32_t y[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t src = w[k >> 1];
                    const float lo = __fadd_rn(__uint_as_float(prmt(src, 0x4B000000u, (k & 1) ? 0x7652u : 0x7650u)), -8388736.0f);
                    const float hi = __fadd_rn(__uint_as_float(prmt(src, 0x4B000000u, (k & 1) ? 0x7653u : 0x7651u)), -8388736.0f);
                    uint32_t d = cvt_bf16x2(hi, lo);
                    if (ZP) d = hsub2(d, z2);
                    y[k] = hmul2(d, s2);
                }
                stg_stream((uint4*)p.out + row * p.cpr + c, make_uint4(y[0], y[1], y[2], y[3]));
            }
        }
    }
}

template <int QT, int OP, bool ZP>
int launch_variant(const EwFast& p, cudaStream_t st) {
    static const int per = [] {
        int n = 0;
        return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, elementwise_fast_kernel<QT, OP, ZP, 4>, 256, 0) == cudaSuccess && n > 0) ? n : 4;
    }();
    const unsigned grid = (unsigned)max((int64_t)1, min((int64_t)kNumSMs * per, (p.rows + 7) / 8));  // one wave of the kernel's residency
    elementwise_fast_kernel<QT, OP, ZP, 4><<<grid, 256, 0, st>>>(p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// qparam geometry of the strategy; false when a chunk of 8 could straddle two qparams or the widths are not powers of two
static bool ew_geometry(const ElemParams& e, EwFast& p) {
    p.x = (const uint4*)e.x; p.scale = (const uint16_t*)e.scale; p.zp = e.zp; p.out = e.out; p.rows = e.rows; p.cpr = (int)(e.cols >> 3);
    auto log2_chunks = [](int64_t width) -> int {  // width % 8 == 0 and width / 8 a power of two, else -1
        if (width <= 0 || width % 8 != 0) return -1;
        const int64_t c = width / 8;
        return (c & (c - 1)) == 0 ? __builtin_ctzll((unsigned long long)c) : -1;
    };
    if (e.strategy == ST_TENSOR) { p.sh = 31; p.rows_per_q = e.rows; p.q_row_stride = 0; }
    else if (e.strategy == ST_CHANNEL) { p.sh = 31; p.rows_per_q = 1; p.q_row_stride = 1; }
    else if (e.strategy == ST_GROUP) {
        if (e.cols < e.group) { p.sh = 31; p.rows_per_q = 1; p.q_row_stride = 1; }  // CT: a row shorter than the group is one group
        else {
            p.sh = log2_chunks(e.group);
            if (p.sh < 0 || e.cols % e.group != 0) return false;
            p.rows_per_q = 1; p.q_row_stride = e.cols / e.group;
        }
    } else if (e.strategy == ST_BLOCK) {
        p.sh = log2_chunks(e.bw);
        if (p.sh < 0 || e.bh <= 0) return false;
        p.rows_per_q = e.bh; p.q_row_stride = (e.cols + e.bw - 1) / e.bw;
    } else return false;
    return true;
}

}  // namespace

// bf16, no global scale; INT4 (codes as int8) or FP8; op EW_QUANT / EW_FQ.  B200Q_ENOSYS when the scheme / shape is not covered.
int launch_elementwise_fast(int op, int qt, const ElemParams& e, cudaStream_t st) {
    if (op != EW_QUANT && op != EW_FQ && op != EW_PACK) return B200Q_ENOSYS;
    if (op == EW_PACK && qt != QT_INT) return B200Q_ENOSYS;
    if (e.gs != nullptr || e.cols % 8 != 0 || e.cols >= (1ll << 33) || e.rows * e.cols == 0) return B200Q_ENOSYS;
    if ((((uintptr_t)e.x) & 15) != 0 || (((uintptr_t)e.out) & 15) != 0 || (((uintptr_t)e.scale) & 1) != 0) return B200Q_ENOSYS;
    if (!(qt == QT_FP8 || (qt == QT_INT && e.nbits == 4))) return B200Q_ENOSYS;
    EwFast p{};
    if (!ew_geometry(e, p)) return B200Q_ENOSYS;
    const bool zp = qt == QT_INT ? e.zp != nullptr : e.has_zp != 0;
    if (qt == QT_INT) {
        if (op == EW_PACK) return zp ? launch_variant<QT_INT, EW_PACK, true>(p, st) : launch_variant<QT_INT, EW_PACK, false>(p, st);
        if (op == EW_QUANT) return zp ? launch_variant<QT_INT, EW_QUANT, true>(p, st) : launch_variant<QT_INT, EW_QUANT, false>(p, st);
        return zp ? launch_variant<QT_INT, EW_FQ, true>(p, st) : launch_variant<QT_INT, EW_FQ, false>(p, st);
    }
    if (op == EW_QUANT) return zp ? launch_variant<QT_FP8, EW_QUANT, true>(p, st) : launch_variant<QT_FP8, EW_QUANT, false>(p, st);
    return zp ? launch_variant<QT_FP8, EW_FQ, true>(p, st) : launch_variant<QT_FP8, EW_FQ, false>(p, st);
}

// int8 codes (any num_bits <= 8) -> bf16, scale bf16, optional int8 zero point.  B200Q_ENOSYS when not covered.
int launch_dequant_int8_fast(const ElemParams& e, cudaStream_t st) {
    if (e.gs != nullptr || e.cols % 8 != 0 || e.cols >= (1ll << 33) || e.rows * e.cols == 0) return B200Q_ENOSYS;
    if ((((uintptr_t)e.x) & 7) != 0 || (((uintptr_t)e.out) & 15) != 0 || (((uintptr_t)e.scale) & 1) != 0) return B200Q_ENOSYS;
    EwFast p{};
    if (!ew_geometry(e, p)) return B200Q_ENOSYS;
    const unsigned grid = (unsigned)max((int64_t)1, min((int64_t)kNumSMs * 4, (p.rows + 7) / 8));
    if (e.zp) dequant_int8_fast_kernel<true, 4><<<grid, 256, 0, st>>>(p);
    else dequant_int8_fast_kernel<false, 4><<<grid, 256, 0, st>>>(p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace b200q
