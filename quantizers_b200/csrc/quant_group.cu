// quant_group.cu -- GROUP / TENSOR_GROUP strategy kernels (group_size in {16,32,64,128,256}).
//
// One warp owns one row segment of 4 x 256 columns; each lane holds four 8-element chunks (4 x 128-bit
// streaming loads in flight), a quantization group is 2/4/8/16/32 consecutive lanes of one chunk and is
// reduced with xor-shuffles.  A CTA is 8 warps = 8 consecutive rows of the same column tile so that the
// asymmetric zero-points can be nibble-packed along rows (CT packs them with packed_dim=0) inside the kernel.
//
// MODE_COMPRESS : observer -> qparams -> codes -> pack   (LLMC update_weight_zp_scale + CT Compressor.compress)
// MODE_QUANT    : caller-supplied qparams -> un-packed codes (CT quantize, forward.py:37-73)
// MODE_QUANT_PACK: caller-supplied qparams -> packed storage (CT Compressor.compress with the module's qparams)
// MODE_FQ       : caller-supplied qparams -> fake-quantized T  (CT fake_quantize, forward.py:149-181)
// MODE_OBS_FQ   : (optional per-column AWQ scale) -> observer -> qparams -> fake-quantize (-> / scale)
//                 (LLMC AWQModifier._compute_best_scale inner step)
#include "common.cuh"
#include "kernels.cuh"

namespace b200q {

constexpr int U = 4;

__device__ __forceinline__ uint32_t pack8_nibbles(const int c[8]) {
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) w |= ((uint32_t)(c[i] + 8) & 0xfu) << (4 * i);
    return w;
}

template <int DT, int QT, int MODE>
__global__ void __launch_bounds__(256) group_kernel(const GroupParams p) {
    constexpr bool kObserve = MODE == MODE_COMPRESS || MODE == MODE_OBS_FQ;   // qparams derived from the data
    constexpr bool kPacked = MODE == MODE_COMPRESS || MODE == MODE_QUANT_PACK;  // storage layout of the state dict
    constexpr bool kCodes = kPacked || MODE == MODE_QUANT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.y * 8 + warp;
    const int64_t b = blockIdx.z;
    const bool row_ok = row < p.rows;
    const int L = p.group >> 3;  // lanes per group
    const int64_t gtot = p.cols / p.group;
    const int64_t tile_c0 = (int64_t)blockIdx.x * (256 * U);
    const int64_t mat_off = b * p.rows * p.cols;

    __shared__ uint8_t zp_s[8][U * 16];

    Chunk8<DT> ch[U];
#pragma unroll
    for (int j = 0; j < U; j++) {
        const int64_t c0 = tile_c0 + j * 256 + lane * 8;
        // AWQ grid (col_scale_stride != 0): every batch entry is the SAME weight under another ratio's scale vector
        const int64_t in_off = (MODE == MODE_OBS_FQ && p.col_scale_stride != 0) ? 0 : mat_off;
        if (row_ok && c0 < p.cols) load_chunk<DT>(ch[j], p.w, in_off + row * p.cols + c0);
        else zero_chunk<DT>(ch[j]);
    }

    const float lo = (QT == QT_INT) ? -(float)(1 << (p.nbits - 1)) : (QT == QT_FP8 ? -448.0f : -6.0f);
    const float hi = (QT == QT_INT) ? (float)((1 << (p.nbits - 1)) - 1) : (QT == QT_FP8 ? 448.0f : 6.0f);
    // forward_quantize (the AWQ path, MODE_OBS_FQ) always finds a zero-point Parameter on the module -- zeros for symmetric schemes
    // (CT initialize_qparams, force_zero_point = True) -- so `scaled += zp` runs and an exact -0.0 becomes +0.0
    const bool add_zp = (QT == QT_INT) ? (MODE == MODE_OBS_FQ || !p.symmetric) : (p.has_zp != 0);
    float gs = 1.0f;
    if (QT == QT_FP4) gs = p.gs[p.gs_stride ? b : 0];

#pragma unroll
    for (int j = 0; j < U; j++) {
        if (tile_c0 + j * 256 >= p.cols) break;  // warp-uniform
        const int64_t c0 = tile_c0 + j * 256 + lane * 8;
        const bool ok = row_ok && c0 < p.cols;
        const int64_t gi = ok ? c0 / p.group : 0;
        float x[8];
        chunk_to_float<DT>(ch[j], x);

        float cs[8];
        if (MODE == MODE_OBS_FQ && p.col_scale != nullptr) {
            if (ok) {
                const float* csp = p.col_scale + b * p.col_scale_stride;
                const float4 s0 = *reinterpret_cast<const float4*>(csp + c0);
                const float4 s1 = *reinterpret_cast<const float4*>(csp + c0 + 4);
                cs[0] = s0.x; cs[1] = s0.y; cs[2] = s0.z; cs[3] = s0.w; cs[4] = s1.x; cs[5] = s1.y; cs[6] = s1.z; cs[7] = s1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) cs[i] = 1.0f;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = round_to<DT>(fmul(x[i], cs[i]));  // W.mul_(scales.view(1,-1))
        }

        // ---- qparams
        float s = 1.0f, z = 0.0f, s_eff = 1.0f;
        uint8_t scode = 0;
        if (kObserve) {
            if (QT == QT_INT && !p.symmetric) {
                float mn = x[0], mx = x[0];
#pragma unroll
                for (int i = 1; i < 8; i++) { mn = fminf(mn, x[i]); mx = fmaxf(mx, x[i]); }
                mn = subwarp_min(mn, L);
                mx = subwarp_max(mx, L);
                qparams_asym<DT>(mn, mx, lo, hi, s, z);
            } else {
                float a = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; i++) a = fmaxf(a, fabsf(x[i]));
                a = subwarp_max(a, L);
                if (QT == QT_FP4) scode = qparams_fp4<DT>(a, gs, s_eff);
                else s = scale_sym<DT>(a, QT == QT_INT ? (hi - lo) * 0.5f : hi);
            }
        } else if (ok) {
            s = load_T<DT>(p.scale, (b * p.rows + row) * gtot + gi);
            if (QT == QT_INT && p.zp_in != nullptr) z = (float)p.zp_in[(b * p.rows + row) * gtot + gi];
            if (QT == QT_FP4) s_eff = fdiv(s, gs);
        }
        const bool use_zp = kObserve ? add_zp : (QT == QT_INT ? p.zp_in != nullptr : p.has_zp != 0);
        // symmetric INT through the compressor still carries an all-zero zero-point: T(u + 0) == u, skip.

        // ---- qparam outputs
        if (MODE == MODE_COMPRESS) {
            if (ok && (lane % L) == 0) {
                const int64_t si = (b * p.rows + row) * gtot + gi;
                if (QT == QT_FP4) ((uint8_t*)p.scale)[si] = scode;
                else store_T<DT>(p.scale, si, s);
            }
            if (QT == QT_INT && !p.symmetric && (lane % L) == 0) {
                const int off = 1 << (p.nbits - 1);
                zp_s[warp][j * (32 / L) + lane / L] = ok ? (uint8_t)((int)z + off) : (uint8_t)0;
            }
        }
        if (!ok) continue;

        // ---- codes
        const int64_t e0 = ((MODE == MODE_OBS_FQ && p.col_scale_stride != 0) ? b * p.out_batch_stride : mat_off) + row * p.cols + c0;
        if (kCodes) {
            if (QT == QT_INT) {
                int c[8];
#pragma unroll
                for (int i = 0; i < 8; i++) c[i] = quant_int<DT>(x[i], s, z, use_zp, lo, hi);
                if (kPacked && p.nbits == 4) {
                    stg_stream((uint32_t*)p.out + (e0 >> 3), pack8_nibbles(c));
                } else {
                    const int off = kPacked ? 128 : 0;  // pack_to_int32(8 bit): bytes (v+128)
                    uint32_t w0 = 0, w1 = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        w0 |= ((uint32_t)(c[i] + off) & 0xffu) << (8 * i);
                        w1 |= ((uint32_t)(c[i + 4] + off) & 0xffu) << (8 * i);
                    }
                    stg_stream((uint8_t*)p.out + e0, make_uint2(w0, w1));
                }
            } else if (QT == QT_FP8) {
                uint32_t w0 = 0, w1 = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    w0 |= (uint32_t)quant_fp8<DT>(x[i], s, use_zp) << (8 * i);
                    w1 |= (uint32_t)quant_fp8<DT>(x[i + 4], s, use_zp) << (8 * i);
                }
                stg_stream((uint8_t*)p.out + e0, make_uint2(w0, w1));
            } else {
                if (kPacked) {
                    uint32_t w = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) w |= quant_fp4(x[i], s_eff) << (4 * i);
                    stg_stream((uint32_t*)p.out + (e0 >> 3), w);
                } else {  // CT quantize() returns the e2m1 grid values in T
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t n = quant_fp4(x[i], s_eff);
                        const float v = e2m1_value(n & 7u);
                        y[i] = (n & 8u) ? -v : v;
                    }
                    store_chunk_T<DT>(p.out, e0, y);
                }
            }
        } else {  // MODE_FQ / MODE_OBS_FQ
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (QT == QT_INT) y[i] = fq_int<DT>(x[i], s, z, use_zp, lo, hi);
                else if (QT == QT_FP8) y[i] = fq_fp8<DT>(x[i], s, use_zp);
                else y[i] = fq_fp4<DT>(x[i], s_eff);
            }
            if (MODE == MODE_OBS_FQ && p.col_scale != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; i++) y[i] = round_to<DT>(fdiv(y[i], cs[i]));  // / scales.view(1,-1) -> copy into T
            }
            store_chunk_T<DT>(p.out, e0, y);
        }
    }

    if (MODE == MODE_COMPRESS && QT == QT_INT && !p.symmetric) {
        __syncthreads();
        const int pf = 32 / p.nbits;          // zero-points per int32 (8 or 4)
        const int words = 8 / pf;             // words per 8-row block (1 or 2)
        const int gpc = U * (32 / L);         // groups per CTA column tile
        const int64_t zrows = (p.rows + pf - 1) / pf;
        for (int t = threadIdx.x; t < gpc * words; t += blockDim.x) {
            const int g = t % gpc, h = t / gpc;
            const int64_t gcol = tile_c0 / p.group + g;
            const int64_t zr = (int64_t)blockIdx.y * words + h;
            if (gcol < gtot && zr < zrows) {
                uint32_t wv = 0;
                for (int i = 0; i < pf; i++) wv |= (uint32_t)zp_s[h * pf + i][g] << (p.nbits * i);
                p.zp_packed[(b * zrows + zr) * gtot + gcol] = (int32_t)wv;
            }
        }
    }
}

template <int DT, int QT, int MODE>
static int launch_group(const GroupParams& p, int64_t batch, cudaStream_t st) {
    const int64_t n256 = (p.cols + 255) / 256;
    dim3 grid((unsigned)((n256 + U - 1) / U), (unsigned)((p.rows + 7) / 8), (unsigned)batch);
    group_kernel<DT, QT, MODE><<<grid, 256, 0, st>>>(p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

template <int MODE>
int dispatch_group(int dt, int qt, const GroupParams& p, int64_t batch, cudaStream_t st) {
    B200Q_REQUIRE(p.group == 16 || p.group == 32 || p.group == 64 || p.group == 128 || p.group == 256,
                  "group_size %d unsupported (16/32/64/128/256)", p.group);
    B200Q_REQUIRE(p.cols % p.group == 0, "tensor column shape must be divisible by the given group_size %d but got %lld",
                  p.group, (long long)p.cols);
    B200Q_REQUIRE(batch >= 1 && batch <= 65535 && (p.rows + 7) / 8 <= 65535, "batch/rows out of range for one launch");
    B200Q_REQUIRE(((uintptr_t)p.w & 15) == 0, "weight pointer must be 16-byte aligned");
    if (p.rows == 0 || p.cols == 0) return B200Q_OK;
    B200Q_DISPATCH_DT(dt, {
        switch (qt) {
        case QT_INT: return launch_group<DT, QT_INT, MODE>(p, batch, st);
        case QT_FP8: return launch_group<DT, QT_FP8, MODE>(p, batch, st);
        case QT_FP4: return launch_group<DT, QT_FP4, MODE>(p, batch, st);
        default: set_error("bad qtype %d", qt); return B200Q_EINVAL;
        }
    });
    return B200Q_EINVAL;
}

template int dispatch_group<MODE_COMPRESS>(int, int, const GroupParams&, int64_t, cudaStream_t);
template int dispatch_group<MODE_QUANT>(int, int, const GroupParams&, int64_t, cudaStream_t);
template int dispatch_group<MODE_FQ>(int, int, const GroupParams&, int64_t, cudaStream_t);
template int dispatch_group<MODE_OBS_FQ>(int, int, const GroupParams&, int64_t, cudaStream_t);
template int dispatch_group<MODE_QUANT_PACK>(int, int, const GroupParams&, int64_t, cudaStream_t);

}  // namespace b200q
