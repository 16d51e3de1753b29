// awq_fq_fast.cu -- AWQ per-ratio weight update, all grid points of one balance layer in one launch (bf16, INT4 GROUP):
//
//     W'_r = T( fake_quantize( T(W * s_r[None, :]) ) / s_r[None, :] )          r = 0 .. R-1
//
// (LLMC AWQModifier._compute_best_scale inner step: ``W.mul_(scales)`` -> fresh memoryless_minmax observer ->
// CT ``forward_quantize`` -> ``/ scales`` -> copy into the bf16 Parameter)
//
// Round 1 ran this on the generic ``group_kernel<MODE_OBS_FQ>``: two IEEE divisions per element and ratio, ~0.3 of the HBM
// roofline, 4.8 % of a dense layer's search and -- replicated on every rank of the token-sharded MoE mapping -- the whole Amdahl
// term of that leg.  Here:
//   * a warp owns RT rows x 256 columns; the RT weight rows are loaded ONCE into registers and re-used for all R ratios (HBM
//     traffic 2 B read + 2 R B written per element, the algorithmic minimum);
//   * per ratio a lane loads its 8 column scales once and derives their bracketed reciprocals, amortised over the RT rows;
//   * x / s with a bf16 x and a bf16 s is a single reciprocal multiply (exact: see quant_group_tma.cu, ONE); quantize -> clamp ->
//     de-quantize stay in packed bf16x2 (the rounding add 200 trick, clamp in the bf16 domain, (q - z) * s as one HMUL2);
//   * the final division by the fp32 column scale is the bracketed reciprocal (both ends through cvt.rn.bf16x2, IEEE repair of
//     the ~2e-4 ambiguous elements), the only place of the chain where the quotient can sit on a rounding boundary;
//   * the -0.0 that torch.round keeps for (-0.5, -0] survives: sign(result) = sign(pre-round value) whenever the zero point is 0.
// Bit-identical to the generic kernel (tests/test_gpu_awq.py, test_gpu_awq_fast_fq).
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.cuh"

namespace b200q {
namespace {
using namespace fast;

struct AwqFqParams {
    const uint16_t* w;      // bf16 [rows, cols]
    int64_t rows, cols;
    const float* scales;    // fp32 [n_ratios, cols]
    int32_t n_ratios;
    uint16_t* out;          // bf16, ratio r at out + r * out_stride
    int64_t out_stride;
};

__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t bf16x2_of(float v) { return (__float_as_uint(v) >> 16) * 0x10001u; }

// exact scalar chain of one chunk (rare: ambiguous quotient, scale outside the safe range)
template <bool SYM>
__device__ __noinline__ uint4 exact_chunk(const uint4 raw, const float* __restrict__ cs, float s, float z) {  // cs: 8 column scales in global memory
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t o[4];
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        float y[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float x = __uint_as_float(e ? (w[k] & 0xffff0000u) : (w[k] << 16));
            const float c = cs ? cs[2 * k + e] : 1.0f;
            const float xs = round_to<DT_BF16>(fmul(x, c));
            const float fq = fq_int<DT_BF16>(xs, s, z, true, -8.0f, 7.0f);   // zero point always added (0 when symmetric)
            y[e] = round_to<DT_BF16>(fdiv(fq, c));
        }
        o[k] = cvt_bf16x2(y[1], y[0]);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// L = lanes per group (group_size / 8): 4, 8, 16
template <bool SYM, int L>
__global__ void __launch_bounds__(256) awq_fq_grid_kernel(const AwqFqParams p) {
    constexpr int RT = 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * 256 + lane * 8;
    const bool col_ok = c0 < p.cols;
    const int64_t row0 = ((int64_t)blockIdx.y * 8 + warp) * RT;
    if (row0 >= p.rows) return;  // whole warp

    uint4 wr[RT];
#pragma unroll
    for (int i = 0; i < RT; i++) {
        const int64_t row = row0 + i;
        wr[i] = (col_ok && row < p.rows) ? ldg_stream(p.w + row * p.cols + c0) : make_uint4(0, 0, 0, 0);
    }
    const uint32_t k200 = 0x43484348u, k192 = 0x43404340u, k207 = 0x434f434fu;

#pragma unroll 1
    for (int r = 0; r < p.n_ratios; r++) {
        float cs[8];
        if (col_ok) {
            const float4 a = *reinterpret_cast<const float4*>(p.scales + (int64_t)r * p.cols + c0);
            const float4 b = *reinterpret_cast<const float4*>(p.scales + (int64_t)r * p.cols + c0 + 4);
            cs[0] = a.x; cs[1] = a.y; cs[2] = a.z; cs[3] = a.w; cs[4] = b.x; cs[5] = b.y; cs[6] = b.z; cs[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++) cs[e] = 1.0f;
        }
        f32x2 cs2[4], rl2[4], rh2[4];
        bool cs_safe = true;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float a = cs[2 * k], b = cs[2 * k + 1];
            cs_safe = cs_safe && a >= 8.6736173798840355e-19f && a <= 1.152921504606846976e18f && b >= 8.6736173798840355e-19f &&
                      b <= 1.152921504606846976e18f;  // 2^-60 .. 2^60: reciprocal normal, no overflow / flush in the products
            const float ra = rcp_approx(a), rb = rcp_approx(b);
            cs2[k] = pack2(a, b);
            rl2[k] = pack2(__fmul_rn(ra, 0.99999952316284179688f), __fmul_rn(rb, 0.99999952316284179688f));
            rh2[k] = pack2(__fmul_rn(ra, 1.00000047683715820312f), __fmul_rn(rb, 1.00000047683715820312f));
        }
        uint16_t* outr = p.out + (int64_t)r * p.out_stride;

#pragma unroll
        for (int i = 0; i < RT; i++) {   // fully unrolled: wr[i] stays in registers
            const int64_t row = row0 + i;
            if (row >= p.rows) break;  // warp-uniform
            const uint32_t w[4] = {wr[i].x, wr[i].y, wr[i].z, wr[i].w};
            // ---- (1) xs = T(w * s_r)
            uint32_t xs[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float lo, hi;
                unpack2(mul2(bf16x2_to_f32x2_fma(w[k]), cs2[k]), lo, hi);
                xs[k] = cvt_bf16x2(hi, lo);
            }
            // ---- (2) group statistics over L lanes (packed: low half = max, high half = -min for the asymmetric case)
            float s, z = 0.0f;
            if (SYM) {
                uint32_t a = hmaxabs2(hmaxabs2(xs[0], xs[1]), hmaxabs2(xs[2], xs[3]));
                a = hmaxabs2(a, prmt(a, a, 0x1032));
#pragma unroll
                for (int o = 1; o < L; o <<= 1) a = hmaxabs2(a, __shfl_xor_sync(0xffffffffu, a, o));
                const float amax = __uint_as_float((a << 16) & 0x7fff0000u);
                s = div_const_bf16_1(amax, 7.5f);
                if (s == 0.0f) s = eps_of<DT_BF16>();
            } else {
                uint32_t mx = hmax2(hmax2(xs[0], xs[1]), hmax2(xs[2], xs[3]));
                uint32_t mn = hmin2(hmin2(xs[0], xs[1]), hmin2(xs[2], xs[3]));
                mx = hmax2(mx, prmt(mx, mx, 0x1032));
                mn = hmin2(mn, prmt(mn, mn, 0x1032));
                uint32_t pk = prmt(mx, mn ^ 0x80008000u, 0x5410);  // {max, -min}
#pragma unroll
                for (int o = 1; o < L; o <<= 1) pk = hmax2(pk, __shfl_xor_sync(0xffffffffu, pk, o));
                const float fmx = fmaxf(__uint_as_float(pk << 16), 0.0f);
                const float fmn = fminf(-__uint_as_float(pk & 0xffff0000u), 0.0f);
                const float d = round_to<DT_BF16>(__fadd_rn(fmx, -fmn));
                const float s0 = div_const_bf16_1(d, 15.0f);
                float t;
                if (scale_is_safe(__float_as_uint(s0))) t = round_to<DT_BF16>(__fmul_rn(fmn, rcp_approx(s0)));
                else t = round_to<DT_BF16>(__fdiv_rn(fmn, s0));   // 0 / 0 -> NaN -> zero point 0
                z = round_to<DT_BF16>(__fadd_rn(-8.0f, -t));
                z = (z == z) ? rintf(fminf(fmaxf(z, -8.0f), 7.0f)) : 0.0f;
                z = __fadd_rn(z, 0.0f);
                s = s0 == 0.0f ? eps_of<DT_BF16>() : s0;
            }
            const bool unsafe = !scale_is_safe(__float_as_uint(s)) || !cs_safe;
            // ---- (3) quantize, (4) de-quantize, packed bf16x2
            const float rs = rcp_approx(s);
            const f32x2 r2 = pack2(rs, rs);
            const uint32_t s2 = bf16x2_of(s), z2 = bf16x2_of(z);
            const uint32_t nb2 = bf16x2_of(-(200.0f + z));                    // exact in bf16 (integers 192 .. 208)
            const uint32_t gmask = (SYM || z == 0.0f) ? 0x80008000u : 0u;      // zero point 0: the result carries the sign of the pre-round value
            uint32_t o[4], diff = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float lo, hi;
                unpack2(mul2(bf16x2_to_f32x2_fma(xs[k]), r2), lo, hi);
                uint32_t v = cvt_bf16x2(hi, lo);                              // T(x / s)
                v = hadd2(v, z2);                                             // T(+ zp): symmetric modules carry a zero zero-point too
                                                                              // (CT initialize_qparams force_zero_point), so -0.0 -> +0.0
                uint32_t m = hadd2(v, k200);                                  // 200 + RNE(v): ulp 1 in [128, 256)
                m = hmin2(hmax2(m, k192), k207);                              // clamp to [-8, 7]
                const uint32_t t = hadd2(m, nb2) & ~gmask;                    // (q - z), sign cleared when it is restored below
                const uint32_t y = hmul2(t, s2 | (v & gmask));                // T((q - z) * s)
                // ---- (5) T(y / s_r): bracketed reciprocal
                const f32x2 yf = bf16x2_to_f32x2_fma(y);
                float al, ah, bl, bh;
                unpack2(mul2(yf, rl2[k]), al, ah);
                unpack2(mul2(yf, rh2[k]), bl, bh);
                o[k] = cvt_bf16x2(ah, al);
                diff |= o[k] ^ cvt_bf16x2(bh, bl);
            }
            uint4 res = make_uint4(o[0], o[1], o[2], o[3]);
            if (diff != 0 || unsafe) res = exact_chunk<SYM>(wr[i], col_ok ? p.scales + (int64_t)r * p.cols + c0 : nullptr, s, z);
            if (col_ok) stg_stream(outr + row * p.cols + c0, res);
        }
    }
}

template <bool SYM, int L>
int launch_v(const AwqFqParams& p, cudaStream_t st) {
    dim3 grid((unsigned)((p.cols + 255) / 256), (unsigned)((p.rows + 63) / 64));
    awq_fq_grid_kernel<SYM, L><<<grid, 256, 0, st>>>(p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace

// bf16 INT4 GROUP (g32 / g64 / g128), col scales [n_ratios, cols] fp32; B200Q_ENOSYS when not covered
int launch_awq_fq_grid_fast(const GroupParams& gp, int n_ratios, cudaStream_t st) {
    static const bool legacy = getenv("B200Q_AWQ_FQ_LEGACY") != nullptr;
    if (legacy || gp.nbits != 4 || gp.col_scale == nullptr) return B200Q_ENOSYS;
    const int g = gp.group;
    if (!(g == 32 || g == 64 || g == 128) || gp.cols % g != 0 || gp.cols % 8 != 0) return B200Q_ENOSYS;
    if ((((uintptr_t)gp.w) & 15) != 0 || (((uintptr_t)gp.out) & 15) != 0 || (((uintptr_t)gp.col_scale) & 15) != 0) return B200Q_ENOSYS;
    if (gp.rows == 0 || gp.cols == 0) return B200Q_OK;
    if ((gp.rows + 63) / 64 > 65535) return B200Q_ENOSYS;
    AwqFqParams p{};
    p.w = (const uint16_t*)gp.w;
    p.rows = gp.rows;
    p.cols = gp.cols;
    p.scales = gp.col_scale;
    p.n_ratios = n_ratios;
    p.out = (uint16_t*)gp.out;
    p.out_stride = n_ratios > 1 ? gp.out_batch_stride : gp.rows * gp.cols;
    if (n_ratios > 1 && (gp.out_batch_stride % 8) != 0) return B200Q_ENOSYS;
    switch (g) {
    case 32: return gp.symmetric ? launch_v<true, 4>(p, st) : launch_v<false, 4>(p, st);
    case 64: return gp.symmetric ? launch_v<true, 8>(p, st) : launch_v<false, 8>(p, st);
    default: return gp.symmetric ? launch_v<true, 16>(p, st) : launch_v<false, 16>(p, st);
    }
}

}  // namespace b200q
