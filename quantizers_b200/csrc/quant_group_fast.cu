// quant_group_fast.cu -- bf16 GROUP-strategy fused compress (INT4 pack-quantized, FP8 float-quantized), tuned for the
// B200 issue budget.
//
// At 6.5 TB/s one SM must retire ~9 bf16 weights per cycle, i.e. the whole observer -> qparams -> quantize -> pack
// chain has ~13 issue slots per element.  The generic kernel (quant_group.cu) spends ~65 (IEEE divisions, scalar
// fp32 rounding emulation, qparams recomputed by every lane).  This kernel keeps the reference's rounding chain
// bit-exact but restructures it:
//   * group min/max on packed bf16x2 (HMNMX2, .xorsign.abs for the symmetric |.|max), xor-shuffle butterfly;
//   * qparams computed ONCE per group: the reduced statistics are transposed so lane i owns group i of the warp
//     tile, (scale, zero-point) travel back as one 32-bit shuffle word;
//   * T(x / s) through the bracketed reciprocal of fastmath.cuh (packed FMUL2 + F2FP, exact fallback per element);
//   * INT4: + zp as packed HFMA2 (single rounding == the reference's fp32-add-then-round for these operand ranges,
//     verified exhaustively on the CPU), round-half-even by the 2^7*1.5625 "magic add" whose bf16 bit pattern is
//     0x4348 + code, clamp + re-bias in one DPX op (VIADDMNMX.S16x2.RELU), nibbles gathered with PRMT;
//   * FP8: the bf16-rounded quotient is re-expanded and converted with cvt.rn.satfinite.e4m3x2.f32 (the saturation
//     is the reference's clamp to +-448).
// Same CTA shape as the generic kernel (8 warps = 8 rows x 1024 columns) so zero-points pack along rows in smem.
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.cuh"

namespace b200q {

namespace {

using namespace fast;
constexpr int U = 4;

// Patch the elements of one 8-element chunk whose bracket ends disagree (or all of them when the scale is outside
// the fast-path range) with the exact IEEE chain.
__device__ __forceinline__ uint32_t cvt_e4m3x2(float hi, float lo) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}

template <bool SYM>
__device__ __noinline__ uint32_t fix_chunk(const uint4 raw, float s, float z, bool all, uint32_t packed) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    Bracket br;
    br.init(s);
    float rl, rh, dummy;
    unpack2(br.lo, rl, dummy);
    unpack2(br.hi, rh, dummy);
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        const uint32_t half = (e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16);
        const float x = __uint_as_float(half);
        const bool differ = __float2bfloat16_rn(__fmul_rn(x, rl)) != __float2bfloat16_rn(__fmul_rn(x, rh));
        if (all || differ) {
            const int c = quant_int<DT_BF16>(x, s, z, !SYM, -8.0f, 7.0f);
            packed = (packed & ~(0xfu << (4 * e))) | (((uint32_t)(c + 8) & 0xfu) << (4 * e));
        }
    }
    return packed;
}

// FP8 variant: patch bytes of the two output words
__device__ __noinline__ uint2 fix_chunk_fp8(const uint4 raw, float s, bool add_zp, bool all, uint2 packed) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    Bracket br;
    br.init(s);
    float rl, rh, dummy;
    unpack2(br.lo, rl, dummy);
    unpack2(br.hi, rh, dummy);
    uint32_t o[2] = {packed.x, packed.y};
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        const uint32_t half = (e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16);
        const float x = __uint_as_float(half);
        const bool differ = __float2bfloat16_rn(__fmul_rn(x, rl)) != __float2bfloat16_rn(__fmul_rn(x, rh));
        if (all || differ) {
            const uint32_t c = quant_fp8<DT_BF16>(x, s, add_zp);
            o[e >> 2] = (o[e >> 2] & ~(0xffu << (8 * (e & 3)))) | (c << (8 * (e & 3)));
        }
    }
    return make_uint2(o[0], o[1]);
}

template <int QT, bool SYM, int LOG2L>
__global__ void __launch_bounds__(256, 4) group_fast_bf16_kernel(const GroupParams p) {
    constexpr int L = 1 << LOG2L;   // lanes per group
    constexpr int P = 32 / L;       // groups per 256-column chunk
    constexpr int G = 8 * L;        // group size
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.y * 8 + warp;
    const int64_t b = blockIdx.z;
    const bool row_ok = row < p.rows;
    const int64_t gtot = p.cols / G;
    const int64_t tile_c0 = (int64_t)blockIdx.x * (256 * U);
    const int64_t row_off = (b * p.rows + row) * p.cols;

    __shared__ uint8_t zp_s[8][U * P];

    uint4 raw[U];
#pragma unroll
    for (int j = 0; j < U; j++) {
        const int64_t c0 = tile_c0 + j * 256 + lane * 8;
        raw[j] = (row_ok && c0 < p.cols) ? ldg_stream((const char*)p.w + (row_off + c0) * 2) : make_uint4(0, 0, 0, 0);
    }

    // ---- A. group statistics on packed bf16x2, reduced over the L lanes of each group
    uint32_t st[U];
#pragma unroll
    for (int j = 0; j < U; j++) {
        if (SYM) {
            uint32_t m = hmaxabs2(hmaxabs2(raw[j].x, raw[j].y), hmaxabs2(raw[j].z, raw[j].w));
            m = hmaxabs2(m, prmt(m, m, 0x1032));
#pragma unroll
            for (int o = L >> 1; o > 0; o >>= 1) m = hmaxabs2(m, __shfl_xor_sync(0xffffffffu, m, o));
            st[j] = m;  // |.|max in both halves (sign bit meaningless)
        } else {
            const uint32_t mx = hmax2(hmax2(raw[j].x, raw[j].y), hmax2(raw[j].z, raw[j].w));
            const uint32_t nm = hmin2(hmin2(raw[j].x, raw[j].y), hmin2(raw[j].z, raw[j].w)) ^ 0x80008000u;
            uint32_t m = hmax2(prmt(mx, nm, 0x5410), prmt(mx, nm, 0x7632));  // lo: max, hi: -min
#pragma unroll
            for (int o = L >> 1; o > 0; o >>= 1) m = hmax2(m, __shfl_xor_sync(0xffffffffu, m, o));
            st[j] = m;
        }
    }

    // ---- B. transpose: lane i (< U*P) owns group i of this warp tile; qparams once per group
    uint32_t mine;
    {
        const int src = (lane % P) << LOG2L;
        const uint32_t t0 = __shfl_sync(0xffffffffu, st[0], src), t1 = __shfl_sync(0xffffffffu, st[1], src);
        const uint32_t t2 = __shfl_sync(0xffffffffu, st[2], src), t3 = __shfl_sync(0xffffffffu, st[3], src);
        const int sel = (lane / P) & 3;
        mine = sel == 0 ? t0 : (sel == 1 ? t1 : (sel == 2 ? t2 : t3));
    }
    float s, z = 0.0f;
    if (SYM) {
        s = scale_sym<DT_BF16>(__uint_as_float((mine << 16) & 0x7fff0000u), QT == QT_INT ? 7.5f : 448.0f);
    } else {
        const float mx = __uint_as_float(mine << 16), neg_mn = __uint_as_float(mine & 0xffff0000u);
        qparams_asym<DT_BF16>(-neg_mn, mx, -8.0f, 7.0f, s, z);
    }
    const uint32_t s_bits = __float_as_uint(s) & 0xffff0000u;  // s is a bf16 value: low half is zero
    const uint32_t word = s_bits | (__float_as_uint(z) >> 16);
    {
        const int64_t gi = tile_c0 / G + lane;
        const bool own = lane < U * P && gi < gtot;
        if (own && row_ok) ((uint16_t*)p.scale)[(b * p.rows + row) * gtot + gi] = (uint16_t)(s_bits >> 16);
        if (!SYM && lane < U * P) zp_s[warp][lane] = (own && row_ok) ? (uint8_t)((int)z + 8) : (uint8_t)0;
    }

    // ---- C. quantize + pack
    const uint32_t kMagic = 0x43484348u;   // bf16x2 (200, 200): bits of 200 + n are 0x4348 + n
    const uint32_t kUnbias = 0xbcc0bcc0u;  // s16x2 (-0x4340): (0x4348 + n) - 0x4340 = n + 8
#pragma unroll
    for (int j = 0; j < U; j++) {
        if (tile_c0 + j * 256 >= p.cols) break;  // warp-uniform
        const uint32_t wj = __shfl_sync(0xffffffffu, word, j * P + (lane >> LOG2L));
        const int64_t c0 = tile_c0 + j * 256 + lane * 8;
        const bool ok = row_ok && c0 < p.cols;
        const uint32_t sj = wj & 0xffff0000u;
        Bracket br;
        br.init(__uint_as_float(sj));
        const uint32_t z2 = prmt(wj, wj, 0x1010);
        const uint32_t w[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
        const bool unsafe = !scale_is_safe(sj);
        uint32_t diff = 0;
        if (QT == QT_FP8) {
            const bool add_zp = p.has_zp != 0;
            uint32_t h[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const f32x2 x = bf16x2_to_f32x2(w[i]);
                float al, ah, bl, bh;
                // "+ zero_point(0)" of the reference turns -0.0 into +0.0 before the fp8 cast
                unpack2(add_zp ? mul2_plus0(x, br.lo) : mul2(x, br.lo), al, ah);
                unpack2(add_zp ? mul2_plus0(x, br.hi) : mul2(x, br.hi), bl, bh);
                const uint32_t v = cvt_bf16x2(ah, al);
                diff |= v ^ cvt_bf16x2(bh, bl);
                h[i] = cvt_e4m3x2(__uint_as_float(v & 0xffff0000u), __uint_as_float(v << 16));  // satfinite == clamp +-448
            }
            uint2 packed = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
            if (diff != 0 || unsafe) packed = fix_chunk_fp8(raw[j], __uint_as_float(sj), add_zp, unsafe, packed);
            if (ok) stg_stream((uint8_t*)p.out + row_off + c0, packed);
            continue;
        }
        uint32_t n[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const f32x2 x = bf16x2_to_f32x2(w[i]);
            float al, ah, bl, bh;
            unpack2(mul2(x, br.lo), al, ah);
            unpack2(mul2(x, br.hi), bl, bh);
            uint32_t v = cvt_bf16x2(ah, al);
            diff |= v ^ cvt_bf16x2(bh, bl);
            if (!SYM) v = hadd2(v, z2);                                          // T(u + zp): one rounding
            n[i] = __viaddmin_s16x2_relu(hadd2(v, kMagic), kUnbias, 0x000f000fu);  // RNE, clamp [-8,7], + 8
        }
        const uint32_t x01 = prmt(n[0], n[1], 0x6420), x23 = prmt(n[2], n[3], 0x6420);  // one nibble per byte
        uint32_t packed = prmt(x01 | (x01 >> 4), x23 | (x23 >> 4), 0x6420);
        if (diff != 0 || unsafe) packed = fix_chunk<SYM>(raw[j], __uint_as_float(sj), __uint_as_float(wj << 16), unsafe, packed);
        if (ok) stg_stream((uint32_t*)p.out + ((row_off + c0) >> 3), packed);
    }

    if (!SYM) {
        __syncthreads();
        const int64_t zrows = (p.rows + 7) / 8;
        if (threadIdx.x < U * P) {
            const int64_t gcol = tile_c0 / G + threadIdx.x;
            if (gcol < gtot) {
                uint32_t wv = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) wv |= (uint32_t)zp_s[i][threadIdx.x] << (4 * i);
                p.zp_packed[(b * zrows + blockIdx.y) * gtot + gcol] = (int32_t)wv;
            }
        }
    }
}

template <int QT, bool SYM, int LOG2L>
int launch(const GroupParams& p, int64_t batch, cudaStream_t st) {
    const int64_t n256 = (p.cols + 255) / 256;
    dim3 grid((unsigned)((n256 + U - 1) / U), (unsigned)((p.rows + 7) / 8), (unsigned)batch);
    group_fast_bf16_kernel<QT, SYM, LOG2L><<<grid, 256, 0, st>>>(p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace

// returns B200Q_ENOSYS when the shape/scheme is not covered (caller falls back to the generic kernel)
int launch_group_fast(int qt, const GroupParams& p, int64_t batch, cudaStream_t st) {
    if (p.cols % p.group != 0 || (((uintptr_t)p.w) & 15) != 0) return B200Q_ENOSYS;
    if (batch < 1 || batch > 65535 || (p.rows + 7) / 8 > 65535 || p.rows == 0 || p.cols == 0) return B200Q_ENOSYS;
    if (qt == QT_INT && p.nbits == 4) {
        switch (p.group) {
        case 32: return p.symmetric ? launch<QT_INT, true, 2>(p, batch, st) : launch<QT_INT, false, 2>(p, batch, st);
        case 64: return p.symmetric ? launch<QT_INT, true, 3>(p, batch, st) : launch<QT_INT, false, 3>(p, batch, st);
        case 128: return p.symmetric ? launch<QT_INT, true, 4>(p, batch, st) : launch<QT_INT, false, 4>(p, batch, st);
        case 256: return p.symmetric ? launch<QT_INT, true, 5>(p, batch, st) : launch<QT_INT, false, 5>(p, batch, st);
        default: return B200Q_ENOSYS;
        }
    }
    if (qt == QT_FP8) {
        switch (p.group) {
        case 32: return launch<QT_FP8, true, 2>(p, batch, st);
        case 64: return launch<QT_FP8, true, 3>(p, batch, st);
        case 128: return launch<QT_FP8, true, 4>(p, batch, st);
        case 256: return launch<QT_FP8, true, 5>(p, batch, st);
        default: return B200Q_ENOSYS;
        }
    }
    return B200Q_ENOSYS;
}

}  // namespace b200q
