// quant_tile_fast.cu -- issue-tuned bf16 kernels for the two remaining headline formats:
//   * FP8 128x128 BLOCK (FP8_BLOCK preset): one CTA per tile, the tile lives in registers (8 x 128-bit loads in
//     flight per thread) between the |.|max reduction and the quantization -> one HBM read;
//   * NVFP4 (e2m1 codes + e4m3 per-16 scales): a thread owns whole groups of 16 (two 128-bit loads), so there is
//     no shuffle and the per-group scale chain runs once per group.
// Both use the bracketed reciprocal of fastmath.cuh for the reference's divisions and fall back to the exact IEEE
// chain (qmath.cuh) for the rare elements whose bracket ends disagree.
#include <cstdlib>
#include "async.cuh"
#include "common.cuh"
#include "fastmath.cuh"
#include "fp4.cuh"
#include "kernels.cuh"

namespace b200q {
namespace {
using namespace fast;
using namespace fp4;
using async::lds128;
using async::smem_u32;

// ------------------------------------------------------------------------------------------------ FP8 block
constexpr int NC = 8;

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

// ONE (round 2): single evaluation of the quotient -- x and the block scale are both bf16, so x * rcp.approx(s) rounds to the reference's
// T(x / s) (tests/test_exact_reciprocal.py; same argument as group_tma_kernel / channel_fast_kernel); unsafe scales keep the IEEE repair.
template <bool ADD_ZP, bool FMA, bool ONE>
// 4 CTAs/SM (64 registers): 0.80 -> 0.92 of the HBM roofline against 3 CTAs/SM at 75 registers -- the tile sits in registers
// between the |max| and the conversion, so every extra resident CTA is 32 KB more in flight
__global__ void __launch_bounds__(256, 4) block_fp8_fast_kernel(const TileParams p) {
    __shared__ uint32_t sm[8];
    const int64_t b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * 128, c0 = (int64_t)blockIdx.x * 128;
    const int64_t mat = b * p.rows * p.cols;
    uint4 raw[NC];
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const int id = i * 256 + threadIdx.x;
        const int64_t r = r0 + (id >> 4), c = c0 + (int64_t)(id & 15) * 8;
        raw[i] = (r < p.rows && c < p.cols) ? ldg_stream((const char*)p.w + (mat + r * p.cols + c) * 2) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < NC; i++) m = hmaxabs2(m, hmaxabs2(hmaxabs2(raw[i].x, raw[i].y), hmaxabs2(raw[i].z, raw[i].w)));
    m = hmaxabs2(m, prmt(m, m, 0x1032));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = hmaxabs2(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    m = hmaxabs2(hmaxabs2(hmaxabs2(sm[0], sm[1]), hmaxabs2(sm[2], sm[3])), hmaxabs2(hmaxabs2(sm[4], sm[5]), hmaxabs2(sm[6], sm[7])));
    const float s = scale_sym<DT_BF16>(__uint_as_float((m << 16) & 0x7fff0000u), 448.0f);
    const uint32_t s_bits = __float_as_uint(s);
    if (threadIdx.x == 0) ((uint16_t*)p.scale)[(b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = (uint16_t)(s_bits >> 16);
    Bracket br;
    if (ONE) br.init1(s); else br.init(s);
    const bool unsafe = !scale_is_safe(s_bits);
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const int id = i * 256 + threadIdx.x;
        const int64_t r = r0 + (id >> 4), c = c0 + (int64_t)(id & 15) * 8;
        if (!(r < p.rows && c < p.cols)) continue;
        const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
        uint32_t h[4], diff = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const f32x2 x = FMA ? bf16x2_to_f32x2_fma(w[k]) : bf16x2_to_f32x2(w[k]);
            float al, ah, bl, bh;
            unpack2(ADD_ZP ? mul2_plus0(x, br.lo) : mul2(x, br.lo), al, ah);
            const uint32_t v = cvt_bf16x2(ah, al);
            if (!ONE) { unpack2(ADD_ZP ? mul2_plus0(x, br.hi) : mul2(x, br.hi), bl, bh); diff |= v ^ cvt_bf16x2(bh, bl); }
            if (FMA) {  // ALU pipe at 72 % (ncu): unpack through FHFMA instead of LOP3 / IMAD.SHL
                float vl, vh;
                unpack2(bf16x2_to_f32x2_fma(v), vl, vh);
                h[k] = cvt_e4m3x2(vh, vl);
            } else {
                h[k] = cvt_e4m3x2(__uint_as_float(v & 0xffff0000u), __uint_as_float(v << 16));
            }
        }
        uint2 packed = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
        if (diff != 0 || unsafe) packed = fix_chunk_fp8(raw[i], s, ADD_ZP, unsafe, packed);
        stg_stream((uint8_t*)p.out + mat + r * p.cols + c, packed);
    }
}

// ------------------------------------------------------------------------------------------------ FP8 per tensor (bf16)
// CT's "FP8" preset (one static scale per weight).  Two launches over the stack -- the whole-matrix |max| must be known before the
// first code -- so the ceiling is 3 / 5 of the single-read roofline; the generic kernels (IEEE division per element) ran at 0.43.
// Pass 1: packed bf16 |max| with 4 x 128-bit loads in flight per thread, lines tagged evict_last so that the tail of the stack is
// still in L2 when pass 2 starts; pass 2: single-evaluation quotient (x and the scale are both bf16) + e4m3 conversion.
__global__ void __launch_bounds__(256) tensor_absmax_bf16_kernel(const uint4* __restrict__ w, int64_t chunks, float* __restrict__ ws) {
    __shared__ uint32_t sm[8];
    const int64_t b = blockIdx.y;
    const uint4* src = w + b * chunks;
    const uint64_t keep = l2_policy_evict_last();
    uint32_t m = 0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    for (; t + 3 * stride < chunks; t += 4 * stride) {
        const uint4 a = ldg_keep(src + t, keep), c = ldg_keep(src + t + stride, keep), d = ldg_keep(src + t + 2 * stride, keep),
                    e = ldg_keep(src + t + 3 * stride, keep);
        m = hmaxabs2(m, hmaxabs2(hmaxabs2(hmaxabs2(a.x, a.y), hmaxabs2(a.z, a.w)), hmaxabs2(hmaxabs2(c.x, c.y), hmaxabs2(c.z, c.w))));
        m = hmaxabs2(m, hmaxabs2(hmaxabs2(hmaxabs2(d.x, d.y), hmaxabs2(d.z, d.w)), hmaxabs2(hmaxabs2(e.x, e.y), hmaxabs2(e.z, e.w))));
    }
    for (; t < chunks; t += stride) {
        const uint4 a = ldg_keep(src + t, keep);
        m = hmaxabs2(m, hmaxabs2(hmaxabs2(a.x, a.y), hmaxabs2(a.z, a.w)));
    }
    m = hmaxabs2(m, prmt(m, m, 0x1032));
    uint32_t bits = (m << 16) & 0x7fff0000u;                   // non-negative fp32 bits order like uints
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < 8; i++) bits = max(bits, sm[i]);
        atomicMax((unsigned int*)&ws[b], bits);
    }
}
template <bool ADD_ZP>
__global__ void __launch_bounds__(256, 6) tensor_fp8_quant_bf16_kernel(const TileParams p, int64_t chunks) {
    const int64_t b = blockIdx.y;
    const float s = scale_sym<DT_BF16>(p.workspace[b], 448.0f);
    const uint32_t s_bits = __float_as_uint(s);
    if (blockIdx.x == 0 && threadIdx.x == 0) ((uint16_t*)p.scale)[b] = (uint16_t)(s_bits >> 16);
    Bracket br;
    br.init1(s);                                               // single evaluation: see block_fp8_fast_kernel
    const bool unsafe = !scale_is_safe(s_bits);
    const uint4* src = reinterpret_cast<const uint4*>(p.w) + b * chunks;
    uint2* dst = reinterpret_cast<uint2*>(p.out) + b * chunks;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t t0 = (int64_t)blockIdx.x * 256 + threadIdx.x; t0 < chunks; t0 += 2 * stride) {
        uint4 raw[2];
        raw[0] = ldg_stream(src + t0);
        const bool two = t0 + stride < chunks;
        raw[1] = two ? ldg_stream(src + t0 + stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (u == 1 && !two) break;
            const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
            uint32_t h[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float al, ah;
                const f32x2 x = bf16x2_to_f32x2(w[k]);
                unpack2(ADD_ZP ? mul2_plus0(x, br.lo) : mul2(x, br.lo), al, ah);
                const uint32_t v = cvt_bf16x2(ah, al);
                h[k] = cvt_e4m3x2(__uint_as_float(v & 0xffff0000u), __uint_as_float(v << 16));
            }
            uint2 packed = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
            if (unsafe) packed = fix_chunk_fp8(raw[u], s, ADD_ZP, true, packed);
            stg_stream(dst + t0 + u * stride, packed);
        }
    }
}

// ------------------------------------------------------------------------------------------------ NVFP4
// thread = U2 groups of 16 contiguous elements; warp = 512 contiguous columns per step; CTA = 8 rows
constexpr int U2 = 2;

__global__ void __launch_bounds__(256, 4) nvfp4_fast_kernel(const GroupParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.y * 8 + warp;
    const int64_t b = blockIdx.z;
    if (row >= p.rows) return;
    const int64_t gtot = p.cols >> 4;
    const int64_t row_off = (b * p.rows + row) * p.cols;
    const float gs = p.gs[p.gs_stride ? b : 0];
    uint4 raw[U2][2];
    int64_t cbase[U2];
#pragma unroll
    for (int j = 0; j < U2; j++) {
        cbase[j] = ((int64_t)blockIdx.x * U2 + j) * 512 + lane * 16;
        if (cbase[j] < p.cols) {
            const uint4* src = reinterpret_cast<const uint4*>((const char*)p.w + (row_off + cbase[j]) * 2);
            raw[j][0] = __ldg(src);
            raw[j][1] = __ldg(src + 1);
        } else {
            raw[j][0] = raw[j][1] = make_uint4(0, 0, 0, 0);
        }
    }
#pragma unroll
    for (int j = 0; j < U2; j++) {
        if (cbase[j] >= p.cols) continue;
        uint32_t m = hmaxabs2(hmaxabs2(hmaxabs2(raw[j][0].x, raw[j][0].y), hmaxabs2(raw[j][0].z, raw[j][0].w)),
                              hmaxabs2(hmaxabs2(raw[j][1].x, raw[j][1].y), hmaxabs2(raw[j][1].z, raw[j][1].w)));
        m = hmaxabs2(m, prmt(m, m, 0x1032));
        float s_eff;
        const uint8_t code = qparams_fp4<DT_BF16>(__uint_as_float((m << 16) & 0x7fff0000u), gs, s_eff);
        ((uint8_t*)p.scale)[(b * p.rows + row) * gtot + (cbase[j] >> 4)] = code;
        Bracket br;
        br.init(s_eff);
        const bool unsafe = !fp4_scale_is_safe(s_eff);
        uint32_t out[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t w[4] = {raw[j][h].x, raw[j][h].y, raw[j][h].z, raw[j][h].w};
            uint32_t c[4], diff = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const f32x2 x = bf16x2_to_f32x2(w[k]);
                float al, ah, bl, bh;
                unpack2(mul2_plus0(x, br.lo), al, ah);  // + 0.0: exact -0.0 inputs carry no sign nibble (torch.sign(-0.) == 0)
                unpack2(mul2_plus0(x, br.hi), bl, bh);
                c[k] = cvt_e2m1x2(ah, al);              // satfinite == clamp to +-6; sign bit from the pre-round sign
                diff |= c[k] ^ cvt_e2m1x2(bh, bl);
            }
            uint32_t packed = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
            if (diff != 0 || unsafe) packed = fix_group_fp4(raw[j][h], s_eff, packed);
            out[h] = packed;
        }
        stg_stream((uint8_t*)p.out + ((row_off + cbase[j]) >> 1), make_uint2(out[0], out[1]));
    }
}


// ------------------------------------------------------------------------------------------------ NVFP4, flat groups
// [batch, rows, cols] with cols % 16 == 0 is a flat array of 16-element groups.  A thread owns whole groups (2 x LDG.128,
// UF groups in flight); a CTA works inside ONE matrix so that everything that depends only on (e4m3 scale code, global
// scale) is tabulated once per CTA in shared memory: there are just 127 non-negative e4m3 codes, so
//     s_eff = fp32(e4m3) / gs,   the bracketed reciprocal pair of s_eff   and   the "safe" flag
// come from one LDS.128 per group instead of an IEEE division + rcp + 2 multiplies (ncu on the previous kernel: 14.9
// issued instructions per weight, ALU pipe 69 % -- instruction bound at 41 % of HBM).  Per group: 9 packed bf16 max ops,
// the constant division by 6 (bracketed), one F2FP to e4m3, the table fetch; per element pair: 2 unpack ops, 2 FFMA2,
// 2 F2FP(e2m1x2), 1 LOP3.
constexpr int FP4_THREADS = 256;
// groups per thread per tile.  Measured: 2 groups (16 KB tiles, 40 registers, 6 CTAs/SM) beat 4 groups (32 KB, 64 registers, 4 CTAs/SM)
// by 3-8 % on the fused kernel: the |max| CTAs mostly wait on HBM, so resident CTAs matter more than loads in flight per thread
constexpr int FP4_UF = 2;
constexpr int FP4_TILE_GROUPS = FP4_THREADS * FP4_UF;       // 512 groups = 16 KB of bf16 per CTA tile

__device__ __forceinline__ void fp4_build_table(Fp4Entry* table, float gs) {
    if (threadIdx.x < 128) {
        // code 0 (scale rounds to zero) is replaced by 0.125 = code 0x20 (helpers.py:101-126); 0x7f is NaN, never produced
        const uint32_t code = threadIdx.x == 0 ? 0x20u : threadIdx.x;
        const float s_eff = fdiv(e4m3_decode((uint8_t)code), gs);
        const float r = rcp_approx(s_eff);
        Fp4Entry e;
        e.r_lo = __fmul_rn(r, 0.99999952316284179688f);
        e.r_hi = __fmul_rn(r, 1.00000047683715820312f);
        e.s_eff = s_eff;
        e.unsafe = fp4_scale_is_safe(s_eff) ? 0.0f : 1.0f;
        table[threadIdx.x] = e;
    }
}

template <bool KEEP = false>
__device__ __forceinline__ void fp4_load_tile(uint4 (&raw)[FP4_UF][2], const uint4* wbase, int64_t g0, int64_t groups_per_mat, uint64_t policy = 0) {
#pragma unroll
    for (int u = 0; u < FP4_UF; u++) {
        const int64_t g = g0 + u * FP4_THREADS;
        if (g < groups_per_mat) {
            raw[u][0] = KEEP ? ldg_keep(wbase + 2 * g, policy) : ldg_stream(wbase + 2 * g);
            raw[u][1] = KEEP ? ldg_keep(wbase + 2 * g + 1, policy) : ldg_stream(wbase + 2 * g + 1);
        } else {
            raw[u][0] = raw[u][1] = make_uint4(0, 0, 0, 0);
        }
    }
}
__device__ __forceinline__ uint32_t fp4_group_absmax2(const uint4 (&r)[2]) {  // packed bf16x2 |max| of 16 elements
    return hmaxabs2(hmaxabs2(hmaxabs2(r[0].x, r[0].y), hmaxabs2(r[0].z, r[0].w)), hmaxabs2(hmaxabs2(r[1].x, r[1].y), hmaxabs2(r[1].z, r[1].w)));
}

// one 1024-group tile of one matrix, already in registers: scale codes + packed e2m1 out
// `locbase` (fused kernel, round 2): the |max| pass already reduced every group; it leaves T(absmax / 6) as bf16 bits in a scratch
// array so that the ALU-bound compress pass skips the 9 packed max ops + the bracketed /6 per group (~1 of its 3.3 ALU-pipe
// instructions per element) and fetches 2 bytes instead.
template <bool FMA>
__device__ __forceinline__ void fp4_compress_regs(const uint4 (&raw)[FP4_UF][2], const Fp4Entry* table, float gs, uint8_t* sbase, uint2* obase,
                                                  int64_t g0, int64_t groups_per_mat, const uint16_t* locbase = nullptr) {
    uint32_t locbits[FP4_UF];
    if (locbase != nullptr) {
#pragma unroll
        for (int u = 0; u < FP4_UF; u++) {
            const int64_t g = g0 + u * FP4_THREADS;
            locbits[u] = g < groups_per_mat ? (uint32_t)__ldcg(locbase + g) << 16 : 0u;   // written by another CTA of this launch: L2, not L1
        }
    }
#pragma unroll
    for (int u = 0; u < FP4_UF; u++) {
        const int64_t g = g0 + u * FP4_THREADS;
        if (g >= groups_per_mat) continue;
        float loc;
        if (locbase != nullptr) {
            loc = __uint_as_float(locbits[u]);
        } else {
            uint32_t m = fp4_group_absmax2(raw[u]);
            m = hmaxabs2(m, prmt(m, m, 0x1032));
            loc = fp4_loc_scale((m << 16) & 0x7fff0000u);                  // T(absmax / 6)
        }
        const float sf = fminf(__fmul_rn(gs, loc), 448.0f);                // gs * loc, clamp (non-negative)
        uint32_t code = cvt_e4m3x2(0.0f, sf) & 0xffu;
        const Fp4Entry e = table[code];
        if (code == 0) code = 0x20;
        sbase[g] = (uint8_t)code;
        const f32x2 rl = pack2(e.r_lo, e.r_lo), rh = pack2(e.r_hi, e.r_hi);
        uint32_t out[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // x * r + 0.0: exact -0.0 inputs carry no sign nibble (torch.sign(-0.) == 0); satfinite == clamp to +-6; the sign
            // bit of a value that rounds to zero comes from the pre-round sign, like the reference's sign(x) * |q|
            const f32x2 x0 = FMA ? bf16x2_to_f32x2_fma(raw[u][h].x) : bf16x2_to_f32x2(raw[u][h].x);
            const f32x2 x1 = FMA ? bf16x2_to_f32x2_fma(raw[u][h].y) : bf16x2_to_f32x2(raw[u][h].y);
            const f32x2 x2 = FMA ? bf16x2_to_f32x2_fma(raw[u][h].z) : bf16x2_to_f32x2(raw[u][h].z);
            const f32x2 x3 = FMA ? bf16x2_to_f32x2_fma(raw[u][h].w) : bf16x2_to_f32x2(raw[u][h].w);
            uint32_t packed = cvt_e2m1x8(mul2_plus0(x0, rl), mul2_plus0(x1, rl), mul2_plus0(x2, rl), mul2_plus0(x3, rl));
            const uint32_t diff = packed ^ cvt_e2m1x8(mul2_plus0(x0, rh), mul2_plus0(x1, rh), mul2_plus0(x2, rh), mul2_plus0(x3, rh));
            if (diff != 0 || e.unsafe != 0.0f) packed = fix_group_fp4(raw[u][h], e.s_eff, packed);
            out[h] = packed;
        }
        stg_stream(obase + g, make_uint2(out[0], out[1]));
    }
}
template <bool FMA>
__device__ __forceinline__ void fp4_compress_tile(const Fp4Entry* table, float gs, const uint4* wbase, uint8_t* sbase, uint2* obase, int64_t g0,
                                                  int64_t groups_per_mat, const uint16_t* locbase = nullptr) {
    uint4 raw[FP4_UF][2];
    fp4_load_tile(raw, wbase, g0, groups_per_mat);
    fp4_compress_regs<FMA>(raw, table, gs, sbase, obase, g0, groups_per_mat, locbase);
}

// caller-supplied global scales (fused q/k/v siblings, decompress round trips): one pass
__global__ void __launch_bounds__(FP4_THREADS, 6) nvfp4_flat_kernel(const GroupParams p, int64_t groups_per_mat, int tiles_per_mat) {
    __shared__ Fp4Entry table[128];
    const int64_t b = blockIdx.y;
    const float gs = p.gs[p.gs_stride ? b : 0];
    fp4_build_table(table, gs);
    __syncthreads();
    const uint4* wbase = reinterpret_cast<const uint4*>((const char*)p.w + b * groups_per_mat * 32);
    uint8_t* sbase = (uint8_t*)p.scale + b * groups_per_mat;
    uint2* obase = reinterpret_cast<uint2*>((uint8_t*)p.out + b * groups_per_mat * 8);
    for (int tile = blockIdx.x; tile < tiles_per_mat; tile += gridDim.x)
        fp4_compress_tile<false>(table, gs, wbase, sbase, obase, (int64_t)tile * FP4_TILE_GROUPS + threadIdx.x, groups_per_mat);
}

// Caller-supplied group scales AND global scale (nvfp4 Compressor.compress with the module's weight_scale / weight_global_scale,
// b200q_quantize_pack): the scale arrives in the weight dtype (CT keeps it as a Parameter of dtype T holding e4m3-representable
// values).  A scale that is a positive finite e4m3 value takes the per-CTA table; anything else (zero, negative, not representable)
// is quantized with the exact IEEE chain on s / gs.
// OP 0: packed e2m1 codes (quantize_pack); 1: CT quantize -- the e2m1 grid VALUES in bf16 (-0.0 for a negative that rounds to zero);
// 2: CT fake_quantize -- bf16(value * (scale / global_scale)) (forward.py:149-181 with forward_helpers.py:255-266 in fp32).
__device__ __forceinline__ void e2m1x2_to_f32(uint32_t one_byte, float& lo, float& hi) {
    asm("{ .reg .b8 t; .reg .b16 u, l, h; .reg .b32 r; cvt.u16.u32 u, %2; cvt.u8.u16 t, u; cvt.rn.f16x2.e2m1x2 r, t; mov.b32 {l, h}, r; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }"
        : "=f"(lo), "=f"(hi) : "r"(one_byte));
}
template <int OP>
__global__ void __launch_bounds__(FP4_THREADS, 6) nvfp4_supplied_kernel(const GroupParams p, int64_t groups_per_mat, int tiles_per_mat) {
    __shared__ Fp4Entry table[128];
    const int64_t b = blockIdx.y;
    const float gs = p.gs[p.gs_stride ? b : 0];
    fp4_build_table(table, gs);
    __syncthreads();
    const uint4* wbase = reinterpret_cast<const uint4*>((const char*)p.w + b * groups_per_mat * 32);
    const uint16_t* sbase = (const uint16_t*)p.scale + b * groups_per_mat;
    uint2* obase = reinterpret_cast<uint2*>((uint8_t*)p.out + b * groups_per_mat * 8);
    for (int tile = blockIdx.x; tile < tiles_per_mat; tile += gridDim.x) {
        const int64_t g0 = (int64_t)tile * FP4_TILE_GROUPS + threadIdx.x;
        uint4 raw[FP4_UF][2];
        uint32_t sbits[FP4_UF];
        fp4_load_tile(raw, wbase, g0, groups_per_mat);
#pragma unroll
        for (int u = 0; u < FP4_UF; u++) {
            const int64_t g = g0 + u * FP4_THREADS;
            sbits[u] = g < groups_per_mat ? (uint32_t)__ldg(sbase + g) << 16 : 0x3f800000u;
        }
#pragma unroll
        for (int u = 0; u < FP4_UF; u++) {
            const int64_t g = g0 + u * FP4_THREADS;
            if (g >= groups_per_mat) continue;
            const float sf = __uint_as_float(sbits[u]);
            const uint32_t code = cvt_e4m3x2(0.0f, sf) & 0xffu;
            const bool tabulated = code >= 1u && code <= 0x7eu && e4m3_decode((uint8_t)code) == sf;
            Fp4Entry e = table[tabulated ? code : 1u];
            if (!tabulated) {  // the table's recipe (fp4_build_table) on this scale
                e.s_eff = fdiv(sf, gs);
                const float r = rcp_approx(e.s_eff);
                e.r_lo = __fmul_rn(r, 0.99999952316284179688f);
                e.r_hi = __fmul_rn(r, 1.00000047683715820312f);
                e.unsafe = fp4_scale_is_safe(e.s_eff) ? 0.0f : 1.0f;
            }
            const f32x2 rl = pack2(e.r_lo, e.r_lo), rh = pack2(e.r_hi, e.r_hi);
            uint32_t out[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const f32x2 x0 = bf16x2_to_f32x2(raw[u][h].x), x1 = bf16x2_to_f32x2(raw[u][h].y);
                const f32x2 x2 = bf16x2_to_f32x2(raw[u][h].z), x3 = bf16x2_to_f32x2(raw[u][h].w);
                uint32_t packed = cvt_e2m1x8(mul2_plus0(x0, rl), mul2_plus0(x1, rl), mul2_plus0(x2, rl), mul2_plus0(x3, rl));
                const uint32_t diff = packed ^ cvt_e2m1x8(mul2_plus0(x0, rh), mul2_plus0(x1, rh), mul2_plus0(x2, rh), mul2_plus0(x3, rh));
                if (diff != 0 || e.unsafe != 0.0f) packed = fix_group_fp4(raw[u][h], e.s_eff, packed);
                out[h] = packed;
                if (OP != 0) {
                    uint32_t y[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float lo, hi;
                        e2m1x2_to_f32((packed >> (8 * k)) & 0xffu, lo, hi);
                        y[k] = OP == 1 ? cvt_bf16x2(hi, lo) : cvt_bf16x2(__fmul_rn(hi, e.s_eff), __fmul_rn(lo, e.s_eff));
                    }
                    stg_stream(reinterpret_cast<uint4*>((uint16_t*)p.out + b * groups_per_mat * 16) + 2 * g + h, make_uint4(y[0], y[1], y[2], y[3]));
                }
            }
            if (OP == 0) stg_stream(obase + g, make_uint2(out[0], out[1]));
        }
    }
}

// Global scale computed here: the whole-matrix |max| must be known before the first code is emitted, i.e. two passes over the
// weight.  One launch does both.  Block A(s) = |max| items of sibling span s (matrices that share min(global_scale): gate/up
// of one expert), block B(s) = compress items of span s; an item is FP4_NT consecutive tiles of one matrix.  CTAs are laid out
// in the order
//     A(0) .. A(L-1),  B(0), A(L), B(1), A(L+1), ...,  B(S-L) .. B(S-1)
// so the reduction runs L spans (<= ~14 MB; measured: larger windows start missing in L2) ahead of the compression and B's re-read is served by the L2: HBM sees every weight
// once.  A B-CTA polls the span's completion counter.  Hardware dispatches a 1-D grid in blockIdx order, so in practice the
// counter is already complete; the launch does not DEPEND on that: a B-CTA that waits too long reduces the span itself, so it
// never blocks on a CTA that has not been scheduled.
// sync words (zeroed by the launcher): per span {|max| bits, finished A items}.
struct Fp4FusedParams {
    int64_t groups_per_mat;
    int32_t tiles_per_mat, items_per_mat, span, n_spans, lookahead, nt;
    int32_t by_item;  // alternate the two passes item by item instead of span by span
    uint32_t* sync;
    float* gs_out;  // [batch]
    uint16_t* loc;  // optional scratch, bf16 bits of T(group |max| / 6) for every group [batch * groups_per_mat] (written by the |max| pass)
    // n / d for n < 2^31 as (n * magic) >> shift (Granlund-Montgomery, 31-bit dividends: the magic fits 32 bits): the CTA-index decode
    // runs in every thread of every CTA, and three hardware-emulated divisions were ~0.5 instructions per weight
    uint32_t nblk_magic, nblk_shift, ipm_magic, ipm_shift;
    int32_t force_fallback;  // tests: every compress CTA reduces its span itself (B200Q_FP4_FORCE_FALLBACK=1)
};
__device__ __forceinline__ uint32_t fastdiv31(uint32_t n, uint32_t magic, uint32_t shift) { return (uint32_t)(((uint64_t)n * magic) >> shift); }

// |max| bits of tiles [tile0, tile1) of one matrix, reduced over the CTA (valid in thread 0)
__device__ __forceinline__ uint32_t fp4_absmax_tiles(const uint4* wbase, int tile0, int tile1, int64_t groups_per_mat, uint32_t* s_red,
                                                     uint16_t* locbase = nullptr) {
    uint32_t mm = 0;
    const uint64_t keep = l2_policy_evict_last();  // the compress pass re-reads these lines
    for (int tile = tile0; tile < tile1; tile++) {
        uint4 raw[FP4_UF][2];
        const int64_t g0 = (int64_t)tile * FP4_TILE_GROUPS + threadIdx.x;
        fp4_load_tile<true>(raw, wbase, g0, groups_per_mat, keep);
#pragma unroll
        for (int u = 0; u < FP4_UF; u++) {
            uint32_t m = fp4_group_absmax2(raw[u]);
            mm = hmaxabs2(mm, m);
            if (locbase != nullptr && g0 + u * FP4_THREADS < groups_per_mat) {   // this pass waits on HBM: the ALU work is free here
                m = hmaxabs2(m, prmt(m, m, 0x1032));
                locbase[g0 + u * FP4_THREADS] = (uint16_t)(__float_as_uint(fp4_loc_scale((m << 16) & 0x7fff0000u)) >> 16);
            }
        }
    }
    mm = hmaxabs2(mm, prmt(mm, mm, 0x1032));
    uint32_t bits = (mm << 16) & 0x7fff0000u;                  // |max| as fp32 bits: non-negative floats order like uints
    bits = __reduce_max_sync(0xffffffffu, bits);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < FP4_THREADS / 32; i++) bits = max(bits, s_red[i]);
    }
    return bits;
}

template <bool FMA>
__global__ void __launch_bounds__(FP4_THREADS, 6) nvfp4_fused_kernel(const GroupParams p, const Fp4FusedParams f) {
    __shared__ Fp4Entry table[128];
    __shared__ float s_gs;
    __shared__ int s_need_fallback;
    __shared__ uint32_t s_red[FP4_THREADS / 32];
    const int nblk = f.items_per_mat * f.span;                 // items per block
    const int S = f.n_spans, Lh = f.lookahead;
    // Launch order: |max| items of spans 0 .. Lh-1, then the compress pass of span s alternates with the |max| pass of span s + Lh,
    // then the compress items of the last Lh spans; a compress CTA never finds its span incomplete (all |max| items of span s
    // precede it).  The two passes alternate span by span when several spans fit the look-ahead window (MoE experts) and ITEM BY
    // ITEM when Lh == 1 (dense shapes, spans of 20-100 MB): whole-span alternation would then serialise the HBM-bound |max| pass
    // and the ALU-bound compress pass (0.49-0.50 -> 0.55-0.56 of the roofline on 50-100 MB spans).
    const unsigned head = (unsigned)Lh * (unsigned)nblk, mid = (unsigned)(S - Lh) * 2u * (unsigned)nblk;
    bool is_b;
    int s, j;
    if (blockIdx.x < head) {
        is_b = false; s = (int)(blockIdx.x / (unsigned)nblk); j = (int)(blockIdx.x - (unsigned)s * (unsigned)nblk);
    } else if (blockIdx.x < head + mid) {
        const unsigned b = blockIdx.x - head, pair = b / (2u * (unsigned)nblk), q = b - pair * 2u * (unsigned)nblk;
        if (f.by_item) { is_b = (q & 1u) == 0; j = (int)(q >> 1); }             // B(s)[0], A(s+Lh)[0], B(s)[1], A(s+Lh)[1], ...
        else { is_b = q < (unsigned)nblk; j = (int)(is_b ? q : q - (unsigned)nblk); }  // B(s)[0..], then A(s+Lh)[0..]
        s = is_b ? (int)pair : (int)pair + Lh;
    } else {
        const unsigned b = blockIdx.x - head - mid, t = b / (unsigned)nblk;
        is_b = true; s = S - Lh + (int)t; j = (int)(b - t * (unsigned)nblk);
    }
    const int mi = j / f.items_per_mat, item = j - mi * f.items_per_mat;
    const int64_t m = (int64_t)s * f.span + mi;
    const int tile0 = item * f.nt, tile1 = min(tile0 + f.nt, f.tiles_per_mat);
    uint32_t* st = f.sync + 2 * s;
    const uint4* wbase = reinterpret_cast<const uint4*>((const char*)p.w + m * f.groups_per_mat * 32);
    uint16_t* locbase = f.loc ? f.loc + m * f.groups_per_mat : nullptr;
    if (!is_b) {
        const uint32_t bits = fp4_absmax_tiles(wbase, tile0, tile1, f.groups_per_mat, s_red, locbase);
        if (threadIdx.x == 0) {
            atomicMax(&st[0], bits);
            __threadfence();
            atomicAdd(&st[1], 1u);
        }
        return;
    }
    uint4 raw[FP4_UF][2];
    fp4_load_tile(raw, wbase, (int64_t)tile0 * FP4_TILE_GROUPS + threadIdx.x, f.groups_per_mat);  // in flight while we poll
    if (threadIdx.x == 0) {
        int spins = 0;
        while (ld_acquire(&st[1]) < (uint32_t)nblk && spins < 20000) { __nanosleep(64); spins++; }
        s_need_fallback = ld_acquire(&st[1]) < (uint32_t)nblk;
    }
    __syncthreads();
    uint32_t bits;
    if (s_need_fallback) {  // never taken when CTAs start in blockIdx order; keeps the kernel independent of that assumption
        bits = 0;
        for (int q = 0; q < f.span; q++) {
            const uint4* wq = reinterpret_cast<const uint4*>((const char*)p.w + ((int64_t)s * f.span + q) * f.groups_per_mat * 32);
            bits = max(bits, fp4_absmax_tiles(wq, 0, f.tiles_per_mat, f.groups_per_mat, s_red));   // (does not touch the loc scratch)
        }
    } else {
        bits = threadIdx.x == 0 ? ld_acquire(&st[0]) : 0u;
    }
    if (threadIdx.x == 0) {
        const float gs0 = gparam<DT_BF16>(__uint_as_float(bits));
        s_gs = gs0;
        if (item == 0) f.gs_out[m] = gs0;
    }
    __syncthreads();
    const float gs = s_gs;
    fp4_build_table(table, gs);
    __syncthreads();
    uint8_t* sbase = (uint8_t*)p.scale + m * f.groups_per_mat;
    uint2* obase = reinterpret_cast<uint2*>((uint8_t*)p.out + m * f.groups_per_mat * 8);
    // the fallback reduced the span without the scratch: recompute the group statistics in that (never observed) case
    const uint16_t* loc_in = s_need_fallback ? nullptr : locbase;
    fp4_compress_regs<FMA>(raw, table, gs, sbase, obase, (int64_t)tile0 * FP4_TILE_GROUPS + threadIdx.x, f.groups_per_mat, loc_in);
    for (int tile = tile0 + 1; tile < tile1; tile++)
        fp4_compress_tile<FMA>(table, gs, wbase, sbase, obase, (int64_t)tile * FP4_TILE_GROUPS + threadIdx.x, f.groups_per_mat, loc_in);
}


// ------------------------------------------------------------------------------------------------ fused kernel, round-2 diet
// ncu on nvfp4_fused_kernel (profiles/r2_ncu_kernels.txt): 16 issued instructions per weight at 66 % issue utilisation, of which
// only ~3 are the conversion itself (unpack, two FFMA2 bracket ends, two F2FP per element pair).  The rest was bookkeeping:
// 64-bit group indices and bounds tests per group, spilled base pointers (40-register cap), a run-time "scratch?" branch, the
// bracketed /6 with its range test, a clamp the saturating conversion already performs.  This version keeps the math and drops
// the bookkeeping: per-thread pointers advanced by a constant per tile, one 32-bit "groups left" counter, T(absmax / 6) as a
// single multiply (exhaustively exact over all 32 641 non-negative bf16 values, tests/test_exact_reciprocal.py), no clamp.
__device__ __forceinline__ float fp4_loc_scale1(uint32_t abits) {
    return __uint_as_float(cvt_bf16x2(0.0f, __fmul_rn(__uint_as_float(abits), 1.0f / 6.0f)) << 16);
}

template <bool KEEP>
__device__ __forceinline__ void fp4_load2(uint4 (&raw)[FP4_UF][2], const uint4* wp, int rem, uint64_t policy = 0) {
#pragma unroll
    for (int u = 0; u < FP4_UF; u++) {
        if (u * FP4_THREADS < rem) {
            raw[u][0] = KEEP ? ldg_keep(wp + 2 * u * FP4_THREADS, policy) : ldg_stream(wp + 2 * u * FP4_THREADS);
            raw[u][1] = KEEP ? ldg_keep(wp + 2 * u * FP4_THREADS + 1, policy) : ldg_stream(wp + 2 * u * FP4_THREADS + 1);
        } else {
            raw[u][0] = raw[u][1] = make_uint4(0, 0, 0, 0);
        }
    }
}

// lean table for the fused kernel: the hot path fetches 8 bytes {r_lo, r_hi}; the repair path re-reads s_eff by code.  An entry whose
// scale is outside the fast range gets the bracket (0, +inf): the two ends then differ for every input (0 vs +-6; NaN for a
// zero), so the ordinary "ends differ" test routes the whole group to the exact chain and the hot loop carries no extra flag.
struct Fp4Lean { float2 r[128]; float s_eff[128]; };
__device__ __forceinline__ void fp4_build_lean(Fp4Lean* t, float gs) {
    if (threadIdx.x < 128) {
        const uint32_t code = threadIdx.x == 0 ? 0x20u : threadIdx.x;       // see fp4_build_table
        const float s_eff = fdiv(e4m3_decode((uint8_t)code), gs);
        const float r = rcp_approx(s_eff);
        const bool safe = fp4_scale_is_safe(s_eff);
        t->r[threadIdx.x] = make_float2(safe ? __fmul_rn(r, 0.99999952316284179688f) : 0.0f,
                                        safe ? __fmul_rn(r, 1.00000047683715820312f) : __uint_as_float(0x7f800000u));
        t->s_eff[threadIdx.x] = s_eff;
    }
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t cvt_e4m3_code(float sf) {   // e4m3 code of a non-negative scale, zero-extended
    uint32_t r;
    asm("{ .reg .b16 t; cvt.rn.satfinite.e4m3x2.f32 t, %1, %2; cvt.u32.u16 %0, t; }" : "=r"(r) : "f"(0.0f), "f"(sf));
    return r;
}
// repair path of one 16-element group (both halves): s_eff is re-read from the table by entry address
static __device__ __noinline__ uint2 fix_group16_fp4(const uint4 a, const uint4 b, uint32_t table_smem, uint32_t code, uint2 packed) {
    const float s_eff = lds_f1(table_smem + 1024 + 4 * code);
    packed.x = fix_group_fp4(a, s_eff, packed.x);
    packed.y = fix_group_fp4(b, s_eff, packed.y);
    return packed;
}

// one 16-element group (two 128-bit words of bf16): scale code + packed e2m1 out.
// Why the bracket (two evaluations) stays while the INT4 kernels got away with one: there x and s are both bf16, so x / s is a ratio of
// two 8-bit significands and can never come closer than 2^-17 to a rounding boundary.  Here s_eff = fp32(e4m3 / gs) is an arbitrary fp32
// number; with gs a bf16 value the IDEAL quotient x * gs / c (X, G 8-bit, C 4-bit significands) either misses every e2m1 boundary
// n * 2^j (n in 1, 3, 5, 7) by >= 2^-16 or hits it exactly (X * G == n * C * 2^j -- common: G = 128, or 3 | G with C = 12, ...).  In the
// exact-hit case the reference's fp32(x / s_eff) lands on the boundary or one ulp to either side depending on the rounding error of
// s_eff, i.e. on the table entry AND on n; no single multiplier reproduces all of those outcomes (an exact tie needs |c * r / gs - 1| <
// 2^-25).  The bracket detects exactly these elements (the two ends straddle the boundary) and sends them to the IEEE chain.
template <bool FMA>
__device__ __forceinline__ void fp4_compress_group(const uint4 r0, const uint4 r1, uint32_t table_smem, float gs, uint8_t* sp, uint2* op) {
    uint32_t m = hmaxabs2(hmaxabs2(hmaxabs2(r0.x, r0.y), hmaxabs2(r0.z, r0.w)), hmaxabs2(hmaxabs2(r1.x, r1.y), hmaxabs2(r1.z, r1.w)));
    m = hmaxabs2(m, prmt(m, m, 0x1032));
    // T(absmax / 6): one multiply (fp4_loc_scale1); |.| is an operand modifier, the bf16 result lands in the high half
    const float loc = __uint_as_float(cvt_bf16x2(__fmul_rn(fabsf(__uint_as_float(m << 16)), 1.0f / 6.0f), 0.0f));
    // satfinite is the reference's clamp to +-448 (the product is non-negative)
    const uint32_t code = cvt_e4m3_code(__fmul_rn(gs, loc));
    const float2 rr = lds_f2(table_smem + 8 * code);
    *sp = (uint8_t)(code == 0 ? 0x20u : code);
    const f32x2 rl = pack2(rr.x, rr.x), rh = pack2(rr.y, rr.y);
    uint2 out, other;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint4 r = h ? r1 : r0;
        const f32x2 x0 = FMA ? bf16x2_to_f32x2_fma(r.x) : bf16x2_to_f32x2(r.x);
        const f32x2 x1 = FMA ? bf16x2_to_f32x2_fma(r.y) : bf16x2_to_f32x2(r.y);
        const f32x2 x2 = FMA ? bf16x2_to_f32x2_fma(r.z) : bf16x2_to_f32x2(r.z);
        const f32x2 x3 = FMA ? bf16x2_to_f32x2_fma(r.w) : bf16x2_to_f32x2(r.w);
        (h ? out.y : out.x) = cvt_e2m1x8(mul2_plus0(x0, rl), mul2_plus0(x1, rl), mul2_plus0(x2, rl), mul2_plus0(x3, rl));
        (h ? other.y : other.x) = cvt_e2m1x8(mul2_plus0(x0, rh), mul2_plus0(x1, rh), mul2_plus0(x2, rh), mul2_plus0(x3, rh));
    }
    // one test per group: the two bracket ends agree on all 16 codes (p(differ) ~ 1e-3 per group)
    if (((out.x ^ other.x) | (out.y ^ other.y)) != 0) out = fix_group16_fp4(r0, r1, table_smem, code, out);
    stg_stream(op, out);
}

template <bool FMA, bool FULL>
__device__ __forceinline__ void fp4_compress2(const uint4 (&raw)[FP4_UF][2], uint32_t table_smem, float gs, uint8_t* sp, uint2* op, int rem) {
#pragma unroll
    for (int u = 0; u < FP4_UF; u++) {
        if (!FULL && u * FP4_THREADS >= rem) continue;
        fp4_compress_group<FMA>(raw[u][0], raw[u][1], table_smem, gs, sp + u * FP4_THREADS, op + u * FP4_THREADS);
    }
}

// |max| bits of `ntiles` tiles starting at the thread's pointer (the first `nfull` of them whole), reduced over the CTA (valid in
// thread 0).  This pass only waits on HBM, so what matters is bytes in flight: whole tiles go two at a time (8 x 128 bit per thread
// outstanding).  Measured on the expert stacks (scripts/ab_fp4_v2.py): one tile at a time 0.66-0.73 of the roofline, two drained
// together 0.73-0.77, two "rotating" (a tile's registers refilled as soon as it is reduced) 0.68-0.71 -- kept the second.
__device__ __forceinline__ uint32_t fp4_absmax2(const uint4* wp, int ntiles, int nfull, int rem, uint32_t* s_red) {
    uint32_t mm = 0;
    const uint64_t keep = l2_policy_evict_last();  // the compress pass re-reads these lines
    int k = 0;
    for (; k + 2 <= nfull; k += 2) {
        uint4 a[FP4_UF][2], b[FP4_UF][2];
        fp4_load2<true>(a, wp, FP4_TILE_GROUPS, keep);
        fp4_load2<true>(b, wp + 2 * FP4_TILE_GROUPS, FP4_TILE_GROUPS, keep);
#pragma unroll
        for (int u = 0; u < FP4_UF; u++) mm = hmaxabs2(mm, hmaxabs2(fp4_group_absmax2(a[u]), fp4_group_absmax2(b[u])));
        wp += 4 * FP4_TILE_GROUPS;
        rem -= 2 * FP4_TILE_GROUPS;
    }
    for (; k < ntiles; k++) {
        uint4 raw[FP4_UF][2];
        if (k < nfull) fp4_load2<true>(raw, wp, FP4_TILE_GROUPS, keep);      // CTA-uniform: no per-group predicates on whole tiles
        else fp4_load2<true>(raw, wp, rem, keep);
#pragma unroll
        for (int u = 0; u < FP4_UF; u++) mm = hmaxabs2(mm, fp4_group_absmax2(raw[u]));
        wp += 2 * FP4_TILE_GROUPS;
        rem -= FP4_TILE_GROUPS;
    }
    mm = hmaxabs2(mm, prmt(mm, mm, 0x1032));
    uint32_t bits = (mm << 16) & 0x7fff0000u;                  // |max| as fp32 bits: non-negative floats order like uints
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < FP4_THREADS / 32; i++) bits = max(bits, s_red[i]);
    }
    return bits;
}

template <bool FMA, int MINB>
__global__ void __launch_bounds__(FP4_THREADS, MINB) nvfp4_fused2_kernel(const GroupParams p, const Fp4FusedParams f) {
    __shared__ Fp4Lean table;
    __shared__ float s_gs;
    __shared__ int s_need_fallback;
    __shared__ uint32_t s_red[FP4_THREADS / 32];
    // CTA order: see nvfp4_fused_kernel
    const unsigned nblk = (unsigned)(f.items_per_mat * f.span);
    const int S = f.n_spans, Lh = f.lookahead;
    const unsigned head = (unsigned)Lh * nblk, mid = (unsigned)(S - Lh) * 2u * nblk;
    bool is_b;
    int s;
    unsigned j;
    if (blockIdx.x < head) {
        is_b = false; s = (int)fastdiv31(blockIdx.x, f.nblk_magic, f.nblk_shift); j = blockIdx.x - (unsigned)s * nblk;
    } else if (blockIdx.x < head + mid) {
        const unsigned b = blockIdx.x - head, pair = fastdiv31(b >> 1, f.nblk_magic, f.nblk_shift), q = b - pair * 2u * nblk;
        if (f.by_item) { is_b = (q & 1u) == 0; j = q >> 1; }
        else { is_b = q < nblk; j = is_b ? q : q - nblk; }
        s = is_b ? (int)pair : (int)pair + Lh;
    } else {
        const unsigned b = blockIdx.x - head - mid, t = fastdiv31(b, f.nblk_magic, f.nblk_shift);
        is_b = true; s = S - Lh + (int)t; j = b - t * nblk;
    }
    const unsigned mi = f.span == 1 ? 0u : fastdiv31(j, f.ipm_magic, f.ipm_shift), item = j - mi * (unsigned)f.items_per_mat;
    const int64_t m = (int64_t)s * f.span + mi;
    const int tile0 = (int)item * f.nt, ntiles = min(f.nt, f.tiles_per_mat - tile0);
    const int64_t g_first = (int64_t)tile0 * FP4_TILE_GROUPS;
    const bool last_partial = g_first + (int64_t)ntiles * FP4_TILE_GROUPS > f.groups_per_mat;   // CTA-uniform: only a matrix's last tile
    const int nfull = last_partial ? ntiles - 1 : ntiles;
    // groups left from this thread's first group of the tile (<= nt * 512: fits 32 bits); group u of the tile exists iff u * 256 < rem
    const int rem = (int)min(f.groups_per_mat - g_first, (int64_t)f.nt * FP4_TILE_GROUPS) - (int)threadIdx.x;
    const int64_t g_thread = m * f.groups_per_mat + g_first + threadIdx.x;
    uint32_t* st = f.sync + 2 * s;
    const uint4* wp = reinterpret_cast<const uint4*>(p.w) + 2 * g_thread;
    if (!is_b) {
        const uint32_t bits = fp4_absmax2(wp, ntiles, nfull, rem, s_red);
        if (threadIdx.x == 0) {
            atomicMax(&st[0], bits);
            __threadfence();
            atomicAdd(&st[1], 1u);
        }
        return;
    }
    // the first tile is in flight while we poll
    uint4 raw[FP4_UF][2];
    fp4_load2<false>(raw, wp, rem);
    if (threadIdx.x == 0) {
        int spins = f.force_fallback ? 20000 : 0;
        while (ld_acquire(&st[1]) < nblk && spins < 20000) { __nanosleep(64); spins++; }
        const bool fallback = f.force_fallback || ld_acquire(&st[1]) < nblk;
        s_need_fallback = fallback;
        if (!fallback) {
            const float gs0 = gparam<DT_BF16>(__uint_as_float(ld_acquire(&st[0])));
            s_gs = gs0;
            if (item == 0) f.gs_out[m] = gs0;
        }
    }
    __syncthreads();
    if (s_need_fallback) {  // never taken when CTAs start in blockIdx order; keeps the kernel independent of that assumption
        uint32_t bits = 0;
        for (int q = 0; q < f.span; q++) {
            const uint4* wq = reinterpret_cast<const uint4*>(p.w) + 2 * (((int64_t)s * f.span + q) * f.groups_per_mat + threadIdx.x);
            const int64_t left = f.groups_per_mat - threadIdx.x;
            // whole matrix, in pieces whose group count fits the 32-bit counter
            for (int64_t t0 = 0; t0 < f.tiles_per_mat; t0 += 1 << 16) {
                const int nt2 = (int)min((int64_t)(1 << 16), (int64_t)f.tiles_per_mat - t0);
                const int rem2 = (int)min(left - t0 * FP4_TILE_GROUPS, (int64_t)nt2 * FP4_TILE_GROUPS);
                bits = max(bits, fp4_absmax2(wq + 2 * t0 * FP4_TILE_GROUPS, nt2, 0, rem2, s_red));
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            const float gs0 = gparam<DT_BF16>(__uint_as_float(bits));
            s_gs = gs0;
            if (item == 0) f.gs_out[m] = gs0;
        }
        __syncthreads();
    }
    const float gs = s_gs;
    fp4_build_lean(&table, gs);
    __syncthreads();
    uint32_t tb = smem_u32(&table);
    asm volatile("" : "+r"(tb));   // keep it in a register: rematerialising the shared-window base costs 3 uniform instructions per group
    uint8_t* sp = (uint8_t*)p.scale + g_thread;
    uint2* op = reinterpret_cast<uint2*>(p.out) + g_thread;
    int r = rem;
    // (refilling a group's registers with the next tile's group right after its conversion -- one group's loads always in flight --
    // was tried: ptxas then keeps the tile in local memory at 48 registers, and 64 registers cost more residency than it gained)
    // Latency hiding is left to the 40 resident warps.  Tried and measured slower (scripts/ab_fp4_v2.py, expert stacks, fraction of
    // the 6.55 TB/s copy peak; this loop: 0.72-0.76): issuing the next tile's loads before converting the current one (two register
    // tiles, 64 registers, 4 CTAs/SM: 0.66-0.71); refilling a group's registers right after its conversion (ptxas keeps the tile in
    // local memory at 48 registers); per-thread cp.async double buffering through shared memory (0.47-0.50).
    for (int k = 0; k < nfull; k++) {
        if (k) fp4_load2<false>(raw, wp, FP4_TILE_GROUPS);
        fp4_compress2<FMA, true>(raw, tb, gs, sp, op, r);
        wp += 2 * FP4_TILE_GROUPS; sp += FP4_TILE_GROUPS; op += FP4_TILE_GROUPS; r -= FP4_TILE_GROUPS;
    }
    if (last_partial) {
        if (nfull) fp4_load2<false>(raw, wp, r);
        fp4_compress2<FMA, false>(raw, tb, gs, sp, op, r);
    }
}

// caller-supplied global scales, lean version of nvfp4_flat_kernel: the compress pass of the fused kernel alone
__global__ void __launch_bounds__(FP4_THREADS, 5) nvfp4_flat2_kernel(const GroupParams p, int64_t groups_per_mat, int tiles_per_mat) {
    __shared__ Fp4Lean table;
    const int64_t b = blockIdx.y;
    const float gs = p.gs[p.gs_stride ? b : 0];
    fp4_build_lean(&table, gs);
    __syncthreads();
    uint32_t tb = smem_u32(&table);
    asm volatile("" : "+r"(tb));
    for (int tile = blockIdx.x; tile < tiles_per_mat; tile += gridDim.x) {
        const int64_t g_first = (int64_t)tile * FP4_TILE_GROUPS;
        const int64_t g_thread = b * groups_per_mat + g_first + threadIdx.x;
        const int rem = (int)min(groups_per_mat - g_first, (int64_t)FP4_TILE_GROUPS) - (int)threadIdx.x;
        const uint4* wp = reinterpret_cast<const uint4*>(p.w) + 2 * g_thread;
        uint8_t* sp = (uint8_t*)p.scale + g_thread;
        uint2* op = reinterpret_cast<uint2*>(p.out) + g_thread;
        uint4 raw[FP4_UF][2];
        if (g_first + FP4_TILE_GROUPS <= groups_per_mat) {       // CTA-uniform
            fp4_load2<false>(raw, wp, FP4_TILE_GROUPS);
            fp4_compress2<false, true>(raw, tb, gs, sp, op, rem);
        } else {
            fp4_load2<false>(raw, wp, rem);
            fp4_compress2<false, false>(raw, tb, gs, sp, op, rem);
        }
    }
}

}  // namespace

// tuning knob read from the environment (development sweeps); `dflt` when unset
static int tune_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int launch_block_fp8_fast(const TileParams& p, int64_t batch, cudaStream_t st) {
    if (p.cols % 8 != 0 || (((uintptr_t)p.w) & 15) != 0 || batch * p.rows * p.cols == 0) return B200Q_ENOSYS;
    dim3 grid((unsigned)((p.cols + 127) / 128), (unsigned)((p.rows + 127) / 128), (unsigned)batch);
    if (grid.y > 65535 || grid.z > 65535) return B200Q_ENOSYS;
    static const bool fma = getenv("B200Q_FP8_BLOCK_LEGACY_ALU") == nullptr;  // FHFMA unpack: +1 % measured (ALU pipe at 72 %)
    static const bool bracket = getenv("B200Q_FP8_BLOCK_BRACKET") != nullptr;  // round-1 two-ended quotient (A/B)
#define B200Q_BLK(ZP_, FMA_) do { if (bracket) block_fp8_fast_kernel<ZP_, FMA_, false><<<grid, 256, 0, st>>>(p); \
                                  else block_fp8_fast_kernel<ZP_, FMA_, true><<<grid, 256, 0, st>>>(p); } while (0)
    if (p.has_zp) { if (fma) B200Q_BLK(true, true); else B200Q_BLK(true, false); }
    else { if (fma) B200Q_BLK(false, true); else B200Q_BLK(false, false); }
#undef B200Q_BLK
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// bf16 per-tensor FP8; workspace: one fp32 word per matrix (the |max| bits).  B200Q_ENOSYS -> the generic two-pass kernels
int launch_tensor_fp8_fast(const TileParams& p, int64_t batch, cudaStream_t st) {
    const int64_t numel = p.rows * p.cols;
    if (numel % 8 != 0 || batch * numel == 0 || batch > 65535 || p.workspace == nullptr) return B200Q_ENOSYS;
    if ((((uintptr_t)p.w) & 15) != 0 || (((uintptr_t)p.out) & 7) != 0 || (numel * 2) % 16 != 0) return B200Q_ENOSYS;
    const int64_t chunks = numel / 8;
    cudaMemsetAsync(p.workspace, 0, sizeof(float) * batch, st);
    // whole waves of the machine over the stack; a matrix gets at least one CTA
    const int64_t per_mat = max((int64_t)1, min((chunks + 1023) / 1024, max((int64_t)1, (int64_t)kNumSMs * 8 * 4 / batch)));
    tensor_absmax_bf16_kernel<<<dim3((unsigned)per_mat, (unsigned)batch), 256, 0, st>>>(reinterpret_cast<const uint4*>(p.w), chunks, p.workspace);
    const int64_t per_mat_q = max((int64_t)1, min((chunks + 511) / 512, max((int64_t)1, (int64_t)kNumSMs * 6 * 4 / batch)));
    if (p.has_zp) tensor_fp8_quant_bf16_kernel<true><<<dim3((unsigned)per_mat_q, (unsigned)batch), 256, 0, st>>>(p, chunks);
    else tensor_fp8_quant_bf16_kernel<false><<<dim3((unsigned)per_mat_q, (unsigned)batch), 256, 0, st>>>(p, chunks);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

int launch_nvfp4_fast(const GroupParams& p, int64_t batch, cudaStream_t st) {
    if (p.cols % 16 != 0 || (((uintptr_t)p.w) & 15) != 0 || batch * p.rows * p.cols == 0) return B200Q_ENOSYS;
    const int64_t groups_per_mat = p.rows * (p.cols >> 4);
    // every matrix must start 16-byte aligned in the weight and 8-byte aligned in the packed output: groups_per_mat * 32 / * 8 do
    const int64_t tiles = (groups_per_mat + FP4_TILE_GROUPS - 1) / FP4_TILE_GROUPS;
    if (batch > 65535 || tiles > (1ll << 30) || (((uintptr_t)p.out) & 7) != 0) return B200Q_ENOSYS;
    // enough CTAs to fill the machine without rebuilding the table more often than once per ~32 KB
    // whole waves of the kernel's residency (launch bounds: 6 CTAs per SM), rounded DOWN so that no straggler wave is left
    const int64_t want = (int64_t)kNumSMs * 6;
    int64_t gx = tiles;
    if (batch * tiles > 4 * want) gx = max((int64_t)1, min(tiles, 4 * want / batch));
    static const bool flat_v1 = getenv("B200Q_FP4_FUSED_V1") != nullptr;   // round-1 kernel (A/B)
    if (flat_v1) nvfp4_flat_kernel<<<dim3((unsigned)gx, (unsigned)batch), FP4_THREADS, 0, st>>>(p, groups_per_mat, (int)tiles);
    else nvfp4_flat2_kernel<<<dim3((unsigned)gx, (unsigned)batch), FP4_THREADS, 0, st>>>(p, groups_per_mat, (int)tiles);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// bf16 weight, bf16 group scales [batch, rows, cols / 16], fp32 global scale(s); op 0: packed e2m1 codes, 1: grid values (bf16), 2: fake_quantize (bf16)
int launch_nvfp4_supplied(const GroupParams& p, int64_t batch, cudaStream_t st, int op) {
    if (p.cols % 16 != 0 || (((uintptr_t)p.w) & 15) != 0 || (((uintptr_t)p.scale) & 1) != 0 || batch * p.rows * p.cols == 0) return B200Q_ENOSYS;
    const int64_t groups_per_mat = p.rows * (p.cols >> 4);
    const int64_t tiles = (groups_per_mat + FP4_TILE_GROUPS - 1) / FP4_TILE_GROUPS;
    if (batch > 65535 || tiles > (1ll << 30) || (((uintptr_t)p.out) & 7) != 0 || p.gs == nullptr) return B200Q_ENOSYS;
    // whole waves of the kernel's residency (launch bounds: 6 CTAs per SM), rounded DOWN so that no straggler wave is left
    const int64_t want = (int64_t)kNumSMs * 6;
    int64_t gx = tiles;
    if (batch * tiles > 4 * want) gx = max((int64_t)1, min(tiles, 4 * want / batch));
    if (op != 0 && (((uintptr_t)p.out) & 15) != 0) return B200Q_ENOSYS;
    const dim3 grid((unsigned)gx, (unsigned)batch);
    if (op == 0) nvfp4_supplied_kernel<0><<<grid, FP4_THREADS, 0, st>>>(p, groups_per_mat, (int)tiles);
    else if (op == 1) nvfp4_supplied_kernel<1><<<grid, FP4_THREADS, 0, st>>>(p, groups_per_mat, (int)tiles);
    else nvfp4_supplied_kernel<2><<<grid, FP4_THREADS, 0, st>>>(p, groups_per_mat, (int)tiles);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// fused |max| -> global scale -> compress; `span` consecutive matrices share min(global_scale); sync: >= 8 * batch / span bytes
int launch_nvfp4_fused(const GroupParams& p, int64_t batch, int span, float* gs_out, uint32_t* sync, cudaStream_t st, uint16_t* loc_scratch) {
    if (p.cols % 16 != 0 || (((uintptr_t)p.w) & 15) != 0 || (((uintptr_t)p.out) & 7) != 0 || batch * p.rows * p.cols == 0) return B200Q_ENOSYS;
    if (span < 1 || batch % span != 0) return B200Q_ENOSYS;
    const int64_t groups_per_mat = p.rows * (p.cols >> 4);
    const int64_t tiles = (groups_per_mat + FP4_TILE_GROUPS - 1) / FP4_TILE_GROUPS;
    // tiles per CTA: the CTA prologue (index decode, poll, table build) is ~100 instructions per thread, i.e. 0.8 per weight at 4 tiles;
    // 6 tiles measured best on the expert stacks (scripts/ab_fp4_v2.py), fewer when the launch would not fill two waves
    static const bool v1 = getenv("B200Q_FP4_FUSED_V1") != nullptr;   // round-1 kernel (A/B)
    int nt = tune_env("B200Q_FP4_NT", 0);
    if (nt <= 0) {
        nt = v1 ? 4 : 6;
        while (!v1 && nt > 2 && 2 * batch * ((tiles + nt - 1) / nt) < 2 * (int64_t)kNumSMs * 5) nt = nt == 6 ? 3 : 2;
    }
    const int64_t items = (tiles + nt - 1) / nt;
    const int64_t n_spans = batch / span;
    const int64_t grid = 2 * n_spans * items * span;
    if (tiles >= (1ll << 30) || grid >= (1ll << 31)) return B200Q_ENOSYS;
    Fp4FusedParams f;
    f.groups_per_mat = groups_per_mat;
    f.tiles_per_mat = (int)tiles;
    f.items_per_mat = (int)items;
    f.span = span;
    f.n_spans = (int)n_spans;
    const int64_t span_bytes = groups_per_mat * 32 * span;
    f.nt = nt;
    f.lookahead = (int)max((int64_t)1, min(n_spans, ((int64_t)tune_env("B200Q_FP4_LOOKAHEAD_MB", 40) << 20) / max(span_bytes, (int64_t)1)));
    const int by_item = tune_env("B200Q_FP4_BY_ITEM", -1);
    f.by_item = by_item >= 0 ? by_item : (f.lookahead == 1 ? 1 : 0);
    f.sync = sync;
    f.gs_out = gs_out;
    auto magic31 = [](uint32_t d, uint32_t& magic, uint32_t& shift) {   // floor(n / d) == (n * magic) >> shift for every n < 2^31
        uint32_t l = 0;
        while ((1ull << l) < d) l++;
        shift = 31 + l;
        magic = (uint32_t)(((1ull << shift) + d - 1) / d);
    };
    magic31((uint32_t)(items * span), f.nblk_magic, f.nblk_shift);
    magic31((uint32_t)items, f.ipm_magic, f.ipm_shift);
    f.force_fallback = tune_env("B200Q_FP4_FORCE_FALLBACK", 0);
    static const bool no_loc = getenv("B200Q_FP4_NO_LOC") != nullptr;   // A/B: recompute the group statistics in the compress pass (round 1)
    f.loc = no_loc ? nullptr : loc_scratch;
    cudaMemsetAsync(sync, 0, sizeof(uint32_t) * 2 * n_spans, st);
    static const bool fma = getenv("B200Q_FP4_FMA") != nullptr;  // FHFMA unpack (ALU-pipe relief), A/B switch
    if (v1 || f.loc != nullptr) {   // the opt-in group-statistics scratch (B200Q_FP4_LOC) only exists in the round-1 kernel
        if (fma) nvfp4_fused_kernel<true><<<(unsigned)grid, FP4_THREADS, 0, st>>>(p, f);
        else nvfp4_fused_kernel<false><<<(unsigned)grid, FP4_THREADS, 0, st>>>(p, f);
    } else {
        // resident CTAs per SM the register budget is set for: 5 = 48 registers, no spills, 193 instructions per tile; 6 = 40 registers
        // with spilled pointers, 213 (measured 3-6 % slower); 4 = 64 registers, slower
        static const int minb = tune_env("B200Q_FP4_MINB", 5);
        const unsigned g = (unsigned)grid;
#define B200Q_FP4_LAUNCH(FMA_) do { \
            if (minb == 6) nvfp4_fused2_kernel<FMA_, 6><<<g, FP4_THREADS, 0, st>>>(p, f); \
            else if (minb == 4) nvfp4_fused2_kernel<FMA_, 4><<<g, FP4_THREADS, 0, st>>>(p, f); \
            else nvfp4_fused2_kernel<FMA_, 5><<<g, FP4_THREADS, 0, st>>>(p, f); } while (0)
        if (fma) B200Q_FP4_LAUNCH(true); else B200Q_FP4_LAUNCH(false);
#undef B200Q_FP4_LAUNCH
    }
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace b200q
