// quant_group_tma.cu -- bf16 GROUP / TENSOR_GROUP fused compress, TMA-staged "lane owns a group" design.
//
// Why: with a warp spread across a group (quant_group_fast.cu) the group statistics need a shuffle butterfly, the
// qparams have to be transposed to owner lanes and broadcast back, and every lane re-derives reciprocals per
// chunk; ncu shows ~21 issued instructions per weight against a budget of ~13 at HBM speed.  Here shared memory is
// the transposer:
//   * the weight is a flat array of groups ([batch, rows, cols] is contiguous and groups never straddle rows), a
//     warp tile is 4096 consecutive elements (8 KB) = 16 x 16-byte chunks per lane;
//   * tiles are brought in by the TMA engine (1-D cp.async.bulk + mbarrier, 3 stages per warp, issued by lane 0),
//     so loads are fully asynchronous and independent of occupancy; packed codes leave through a bulk store;
//   * lane l owns groups {l, l+32, ...} of the tile and walks its chunks in a rotated order that is bank-conflict
//     free (LDS.128): statistics are a straight max/min chain (no shuffles), qparams are computed once per group by
//     the lane that uses them, the reciprocal bracket (fastmath.cuh) is set up once per group.
// Arithmetic is the same bit-exact chain as the other kernels (qmath.cuh); elements whose reciprocal bracket is
// ambiguous are recomputed with the IEEE path.  Zero-points of asymmetric INT4 are OR-ed into the row-packed int32
// layout with one atomic per group (the buffer is zeroed by the launcher).
#include <cstdlib>
#include "async.cuh"
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.cuh"

namespace b200q {
namespace {
using namespace fast;

constexpr int kStages = 2;   // default ring depth per warp (template parameter STAGES overrides it for the A/B variants)
// warps per CTA (one CTA per SM): as many as shared memory allows.  ncu: with 8 KB tiles (12 warps = 3 per scheduler) neither the
// ALU nor the FMA pipe is saturated (51 % / 31 %), issue slots are 65 % busy and the stalls are fixed-latency waits -- too few
// warps to hide them.  4 KB tiles double the warps per SM; a group of 128 is then shared by a pair of lanes.
template <int QT, int TILE, int STAGES = kStages, int NWARPS = 0> struct WarpsFor {
    static constexpr int value = NWARPS > 0 ? NWARPS : STAGES == 2 ? (TILE == 4096 ? ((QT == QT_FP8) ? 10 : 12) : ((QT == QT_FP8) ? 20 : 24))
                                             : (STAGES == 3 ? 8 : 6);   // 3 x 8 KB x 8 warps / 4 x 8 KB x 6 warps: deeper ring, fewer warps
};

using namespace async;

__device__ __forceinline__ uint32_t cvt_e4m3x2(float hi, float lo) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_e2m1x2(float hi, float lo) {
    uint16_t r;
    asm("{ .reg .b8 t; cvt.rn.satfinite.e2m1x2.f32 t, %1, %2; cvt.u16.u8 %0, t; }" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ bool fp4_scale_is_safe(float s_eff) { return s_eff >= 7.8886090522101181e-31f && s_eff <= 65536.0f; }

struct TmaParams {
    const uint16_t* w;        // flat bf16
    int64_t n_groups;         // batch * rows * cols / g
    int64_t groups_per_mat;   // rows * cols / g
    int64_t groups_per_row;   // cols / g
    int64_t rows;
    uint8_t* out;             // packed codes, flat
    void* scale;              // bf16 (INT/FP8) or e4m3 bytes (FP4), flat [n_groups]
    int32_t* zp_packed;       // asym INT4, legacy path: OR-ed in place with one atomic per group (buffer zeroed by the launcher)
    int8_t* zp_i8;            // asym INT4: int8 [n_groups] scratch, row-packed into zp_packed by zp_pack_rows_kernel afterwards
    const int8_t* zp_in;      // SUPPLIED mode, asym INT4: int8 [n_groups]
    const float* gs;          // FP4: fp32 [batch] (stride 1) or [1] (stride 0)
    int32_t gs_stride;
    int32_t has_zp;
    uint32_t unbias = 0xbcc0bcc0u;  // -0x4340 per s16 half (bf16 bits of 200 + n -> n + 8); a kernel parameter so that ptxas keeps it out of the immediates
};

// ---- exact per-element repair of one chunk (rare): returns the repaired packed representation
template <int QT, bool SYM>
__device__ __noinline__ uint2 repair_chunk(const uint4 raw, float s, float z, bool add_zp, bool all, uint2 packed) {
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
        bool differ;
        if (QT == QT_FP4) differ = cvt_e2m1x2(0.0f, __fmaf_rn(x, rl, 0.0f)) != cvt_e2m1x2(0.0f, __fmaf_rn(x, rh, 0.0f));
        else differ = __float2bfloat16_rn(__fmul_rn(x, rl)) != __float2bfloat16_rn(__fmul_rn(x, rh));
        if (all || differ) {
            if (QT == QT_INT) {
                const uint32_t c = (uint32_t)(quant_int<DT_BF16>(x, s, z, !SYM, -8.0f, 7.0f) + 8) & 0xfu;
                o[0] = (o[0] & ~(0xfu << (4 * e))) | (c << (4 * e));
            } else if (QT == QT_FP8) {
                const uint32_t c = quant_fp8<DT_BF16>(x, s, add_zp);
                o[e >> 2] = (o[e >> 2] & ~(0xffu << (8 * (e & 3)))) | (c << (8 * (e & 3)));
            } else {
                o[0] = (o[0] & ~(0xfu << (4 * e))) | (quant_fp4(x, s) << (4 * e));
            }
        }
    }
    return make_uint2(o[0], o[1]);
}

// LOG2N: log2(chunks per group): 1 (g16) 2 (g32) 3 (g64) 4 (g128)
// FMA: ALU-pipe relief variants (fastmath.cuh): FHFMA unpack, DPX 3-input max for the asymmetric statistics, bracket agreement
// accumulated with HFMA2, nibble folding with LEA.HI
// SUPPLIED: the caller's qparams (Compressor.compress with an observer's weight_scale / weight_zero_point, b200q_quantize_pack): the
// statistics and qparam stages are replaced by one load per group; nothing but the packed codes is written.
// ONE (INT4 / FP8, i.e. bf16 data divided by a bf16 scale): x and s carry 8-bit significands, so x / s can never come closer to a
// bf16 rounding boundary (a 9-bit odd significand) than 1 / (255 * 511) = 2^-16.9 relative -- brute-forced over every pair of
// significands in tests/test_exact_reciprocal.py.  A product x * rcp(s) is within 2^-22 of the quotient, so it rounds to the same
// bf16 as the reference's fp32 division: ONE evaluation, no bracket, no repair.  Only scales outside [2^-100, 1] (products that
// overflow / flush) still take the IEEE chain, group-wide.  NVFP4 divides by an fp32 quotient and keeps the bracket.
template <int QT, bool SYM, int LOG2N, bool FMA, int TILE, bool SUPPLIED = false, bool ONE = false, bool PF = true, int STAGES = kStages, bool SPIN = false,
          int NWARPS = 0>
__global__ void __launch_bounds__(WarpsFor<QT, TILE, STAGES, NWARPS>::value * 32, 1) group_tma_kernel(const TmaParams p) {
    static_assert(!ONE || QT != QT_FP4, "NVFP4 needs the bracket");
    constexpr int kStages = STAGES;
    constexpr int kWarps = WarpsFor<QT, TILE, STAGES, NWARPS>::value;
    constexpr int kTileBytes = TILE * 2;
    constexpr int N = 1 << LOG2N;          // chunks (8 elements, 16 bytes) per group
    constexpr int G = 8 * N;               // group size
    constexpr int GPT = TILE / G;          // groups per tile
    constexpr int LPG = GPT >= 32 ? 1 : 32 / GPT;   // lanes sharing one group (2 for g128 in a 4 KB tile)
    constexpr int GPL = GPT >= 32 ? GPT / 32 : 1;   // groups per lane per tile
    constexpr int NL = N / LPG;            // chunks of a group per lane
    constexpr int OUT_CHUNK = (QT == QT_FP8) ? 8 : 4;  // output bytes per 8-element chunk
    constexpr int OUT_BYTES = (TILE / 8) * OUT_CHUNK;  // per tile
    constexpr int ROT_SHIFT = (LPG > 1 || LOG2N >= 3) ? 0 : (3 - LOG2N);
    static_assert(LPG <= 2 && NL >= 2, "tile / group geometry");

    extern __shared__ __align__(1024) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* my = smem + (size_t)warp * (kStages * kTileBytes + OUT_BYTES + 256);
    const uint32_t in_base = smem_u32(my);
    const uint32_t out_base = in_base + kStages * kTileBytes;
    const uint32_t bar_base = out_base + OUT_BYTES;

    if ((in_base & 255u) != 0) __trap();  // the XOR swizzle below works on absolute addresses

    const int64_t n_tiles = (p.n_groups + GPT - 1) / GPT;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t wstride = (int64_t)gridDim.x * kWarps;

    if (lane == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(bar_base + 8 * s, 1);
        fence_async_smem();
    }
    __syncwarp();

    auto issue = [&](int64_t tile, int stage) {
        const int64_t g0 = tile * GPT;
        const uint32_t bytes = (uint32_t)(min((int64_t)GPT, p.n_groups - g0) * G * 2);
        mbar_arrive_expect_tx(bar_base + 8 * stage, bytes);
        tma_load_1d(in_base + stage * kTileBytes, p.w + g0 * G, bytes, bar_base + 8 * stage);
    };
    if (lane == 0) {
        for (int s = 0; s < kStages; s++) {
            const int64_t t = gwarp + s * wstride;
            if (t < n_tiles) issue(t, s);
        }
    }

    const uint32_t kMagic = 0x43484348u, kUnbias = 0xbcc0bcc0u;
    const int rot = (lane >> ROT_SHIFT) & (NL - 1);
    // opaque copy of the un-bias constant: ptxas otherwise re-materialises the immediate before every VIADDMNMX (4 extra IMAD.MOV per chunk)
    const uint32_t kUnbiasR = p.unbias;  // == kUnbias, from the constant bank
    int it = 0;
    for (int64_t tile = gwarp; tile < n_tiles; tile += wstride, it++) {
        const int stage = it % kStages;
        const uint32_t parity = (uint32_t)(it / kStages) & 1u;
        const int64_t g0 = tile * GPT;
        const int n_here = (int)min((int64_t)GPT, p.n_groups - g0);
        if (SPIN) mbar_wait_spin(bar_base + 8 * stage, parity);
        else mbar_wait(bar_base + 8 * stage, parity);
        const uint32_t tin = in_base + stage * kTileBytes;

        // (the previous tile's bulk store must have finished READING the output staging buffer before it is overwritten: that wait
        // sits right before this tile's first staging store, after the statistics / qparam stages, so the store engine has that
        // long to drain instead of stalling the warp at the top of the tile)

        // position of the tile's first group (warp-uniform, once per tile): matrix b0, row r0, group-in-row k0
        int64_t b0 = 0;
        uint32_t r0 = 0, k0 = 0;
        if (QT == QT_INT && !SYM && !SUPPLIED && p.zp_i8 == nullptr) {
            b0 = g0 / p.groups_per_mat;
            const int64_t rem0 = g0 - b0 * p.groups_per_mat;
            r0 = (uint32_t)(rem0 / p.groups_per_row);
            k0 = (uint32_t)(rem0 - (int64_t)r0 * p.groups_per_row);
        }
        float gs_tile = 1.0f;
        bool gs_uniform = true;
        if (QT == QT_FP4) {
            if (p.gs_stride == 0) gs_tile = p.gs[0];
            else {
                const int64_t b0 = g0 / p.groups_per_mat, b1 = (g0 + n_here - 1) / p.groups_per_mat;
                gs_uniform = (b0 == b1);
                gs_tile = p.gs[b0];
            }
        }

        // FP4: s_eff = fp32(s) / gs with s = M * 2^E (M in 1..15): RN(M / gs) * 2^E is the same fp32 value, so one
        // IEEE division per lane per TILE (lane = M) replaces one per group; groups fetch their entry by shuffle.
        float tab = 0.0f;
        const bool gs_tab_ok = gs_uniform && gs_tile >= 8.6736173798840355e-19f && gs_tile <= 1.152921504606846976e18f;  // 2^+-60
        if (QT == QT_FP4) tab = __fdiv_rn((float)(lane & 15), gs_tile);

#pragma unroll (GPL >= 4 ? 2 : 1)
        for (int gi = 0; gi < GPL; gi++) {
            const int gl = LPG == 1 ? gi * 32 + lane : lane / LPG;   // group index inside the tile
            const int half = LPG == 1 ? 0 : lane % LPG;                // which part of the group this lane owns
            const bool act = gl < n_here;            // partial last tile: inactive lanes compute on stale smem, store nothing
            const bool owner = act && half == 0;     // writes the group's qparams
            const uint32_t gaddr = tin + (uint32_t)gl * (G * 2) + (uint32_t)half * (NL * 16);
            // ---- A. statistics (two independent chains for ILP)
            uint32_t st_a = 0, st_b = 0;
            if (SUPPLIED) {
            } else if (SYM) {
                uint32_t a0 = 0, a1 = 0;
#pragma unroll
                for (int i = 0; i < NL; i += 2) {
                    const uint4 v0 = lds128(gaddr + (((i ^ rot)) << 4));
                    const uint4 v1 = lds128(gaddr + ((((i + 1) ^ rot)) << 4));
                    a0 = hmaxabs2(a0, hmaxabs2(hmaxabs2(v0.x, v0.y), hmaxabs2(v0.z, v0.w)));
                    a1 = hmaxabs2(a1, hmaxabs2(hmaxabs2(v1.x, v1.y), hmaxabs2(v1.z, v1.w)));
                }
                st_a = hmaxabs2(a0, a1);
                st_a = hmaxabs2(st_a, prmt(st_a, st_a, 0x1032));
            } else if (FMA) {
                // max(max, 0) and min(min, 0) are all the asymmetric qparams need (helpers.py:72-73), and both are MAX reductions on
                // the raw bf16 bits: as s16 the largest positive float wins (RELU clamps an all-negative group to +0), as u16 the
                // negative float of largest magnitude wins (a result below 0x8000 means "no negative element").  DPX 3-input
                // VIMNMX3 halves the op count of the HMNMX2 max + min chains.
                uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
#pragma unroll
                for (int i = 0; i < NL; i += 2) {
                    const uint4 v0 = lds128(gaddr + (((i ^ rot)) << 4));
                    const uint4 v1 = lds128(gaddr + ((((i + 1) ^ rot)) << 4));
                    a0 = __vimax3_s16x2_relu(__vimax3_s16x2_relu(v0.x, v0.y, v0.z), v0.w, a0);
                    b0 = __vimax3_u16x2(__vimax3_u16x2(v0.x, v0.y, v0.z), v0.w, b0);
                    a1 = __vimax3_s16x2_relu(__vimax3_s16x2_relu(v1.x, v1.y, v1.z), v1.w, a1);
                    b1 = __vimax3_u16x2(__vimax3_u16x2(v1.x, v1.y, v1.z), v1.w, b1);
                }
                st_a = __vimax3_s16x2_relu(a0, a1, prmt(a0, a1, 0x1032));   // low half: max over both halves of a0 and a1's high
                st_a = __vmaxs2(st_a, prmt(st_a, st_a, 0x1032));
                st_b = __vimax3_u16x2(b0, b1, prmt(b0, b1, 0x1032));
                st_b = __vmaxu2(st_b, prmt(st_b, st_b, 0x1032));
            } else {
                uint32_t a0 = 0xff80ff80u, a1 = 0xff80ff80u, b0 = 0x7f807f80u, b1 = 0x7f807f80u;  // -inf / +inf
#pragma unroll
                for (int i = 0; i < NL; i += 2) {
                    const uint4 v0 = lds128(gaddr + (((i ^ rot)) << 4));
                    const uint4 v1 = lds128(gaddr + ((((i + 1) ^ rot)) << 4));
                    a0 = hmax2(a0, hmax2(hmax2(v0.x, v0.y), hmax2(v0.z, v0.w)));
                    b0 = hmin2(b0, hmin2(hmin2(v0.x, v0.y), hmin2(v0.z, v0.w)));
                    a1 = hmax2(a1, hmax2(hmax2(v1.x, v1.y), hmax2(v1.z, v1.w)));
                    b1 = hmin2(b1, hmin2(hmin2(v1.x, v1.y), hmin2(v1.z, v1.w)));
                }
                st_a = hmax2(a0, a1);
                st_b = hmin2(b0, b1);
                st_a = hmax2(st_a, prmt(st_a, st_a, 0x1032));
                st_b = hmin2(st_b, prmt(st_b, st_b, 0x1032));
            }
            if (LPG == 2 && !SUPPLIED) {  // the other half of the group lives in the neighbouring lane
                const uint32_t oa = __shfl_xor_sync(0xffffffffu, st_a, 1), ob = __shfl_xor_sync(0xffffffffu, st_b, 1);
                if (SYM) st_a = hmaxabs2(st_a, oa);
                else if (FMA) { st_a = __vmaxs2(st_a, oa); st_b = __vmaxu2(st_b, ob); }
                else { st_a = hmax2(st_a, oa); st_b = hmin2(st_b, ob); }
            }
            if (!SYM && FMA && (st_b & 0x8000u) == 0) st_b = 0;               // no negative element: min(min, 0) = 0
            // ---- B. qparams (once per group, by the lanes that use them; same rounding chain as qmath.cuh)
            float s, z = 0.0f;
            Bracket br;
            const int64_t gidx = g0 + gl;
            if (SUPPLIED) {
                const int64_t gq = act ? gidx : g0;  // inactive lanes of a partial tile read a valid entry and store nothing
                if (QT == QT_FP4) s = __fdiv_rn(e4m3_decode(((const uint8_t*)p.scale)[gq]), p.gs[p.gs_stride ? gq / p.groups_per_mat : 0]);
                else s = __uint_as_float((uint32_t)((const uint16_t*)p.scale)[gq] << 16);
                if (QT == QT_INT && !SYM) z = (float)p.zp_in[gq];
                if (ONE) br.init1(s); else br.init(s);
            } else if (QT == QT_FP4) {
                const float amax = __uint_as_float((st_a << 16) & 0x7fff0000u);
                float gsv = gs_tile;
                if (!gs_uniform && act) gsv = p.gs[gidx / p.groups_per_mat];
                const float loc = div_const_bf16(amax, 6.0f);
                float sf = fminf(__fmul_rn(gsv, loc), 448.0f);  // >= 0; helpers.py:101-108
                uint32_t code = cvt_e4m3x2(0.0f, sf) & 0xffu;
                if (code == 0) code = 0x20;                      // zero scale -> 0.125 (helpers.py:364-366)
                const uint32_t e = code >> 3, mm = code & 7u;
                const uint32_t M = e ? (mm | 8u) : mm;
                const int E = e ? (int)e - 10 : -9;
                const float t = __shfl_sync(0xffffffffu, tab, (int)M);
                if (gs_tab_ok) s = __fmul_rn(t, __uint_as_float((uint32_t)(E + 127) << 23));
                else s = __fdiv_rn(e4m3_decode((uint8_t)code), gsv);
                if (owner) ((uint8_t*)p.scale)[gidx] = (uint8_t)code;
                br.init(s);
            } else if (SYM) {
                const float amax = __uint_as_float((st_a << 16) & 0x7fff0000u);
                s = ONE ? div_const_bf16_1(amax, QT == QT_INT ? 7.5f : 448.0f) : div_const_bf16(amax, QT == QT_INT ? 7.5f : 448.0f);
                if (s == 0.0f) s = eps_of<DT_BF16>();
                if (ONE) br.init1(s); else br.init(s);
            } else {
                const float mn = fminf(__uint_as_float(st_b << 16), 0.0f), mx = fmaxf(__uint_as_float(st_a << 16), 0.0f);
                const float d = round_to<DT_BF16>(__fadd_rn(mx, -mn));
                const float s0 = ONE ? div_const_bf16_1(d, 15.0f) : div_const_bf16(d, 15.0f);  // helpers.py:96
                if (ONE) br.init1(s0); else br.init(s0);
                // zp = clamp(T(-8 - T(mn / s0))), helpers.py:97-98 (un-eps'ed scale: 0/0 -> NaN -> 0 after the int8 cast)
                float t;
                if (ONE) {
                    float r1, dummy;
                    unpack2(br.lo, r1, dummy);
                    if (scale_is_safe(__float_as_uint(s0))) t = round_to<DT_BF16>(__fmul_rn(mn, r1));
                    else t = round_to<DT_BF16>(__fdiv_rn(mn, s0));
                } else {
                    float rl, rh, dummy;
                    unpack2(br.lo, rl, dummy);
                    unpack2(br.hi, rh, dummy);
                    const uint32_t u = cvt_bf16x2(__fmul_rn(mn, rh), __fmul_rn(mn, rl));
                    if (((u >> 16) == (u & 0xffffu)) && scale_is_safe(__float_as_uint(s0))) t = __uint_as_float(u << 16);
                    else t = round_to<DT_BF16>(__fdiv_rn(mn, s0));
                }
                z = round_to<DT_BF16>(__fadd_rn(-8.0f, -t));
                z = (z == z) ? rintf(fminf(fmaxf(z, -8.0f), 7.0f)) : 0.0f;
                s = s0;
                if (s0 == 0.0f) { s = eps_of<DT_BF16>(); if (ONE) br.init1(s); else br.init(s); }
            }
            if (QT != QT_FP4 && !SUPPLIED) {
                if (owner) ((uint16_t*)p.scale)[gidx] = (uint16_t)(__float_as_uint(s) >> 16);
                if (QT == QT_INT && !SYM && owner && p.zp_i8 != nullptr) {
                    p.zp_i8[gidx] = (int8_t)(int)z;
                } else if (QT == QT_INT && !SYM && owner) {
                    const uint32_t gpr = (uint32_t)p.groups_per_row;  // launcher guarantees < 2^31
                    const uint32_t t = k0 + (uint32_t)gl, dr = t / gpr, k = t - dr * gpr;
                    int64_t b = b0, r = (int64_t)r0 + dr;
                    while (r >= p.rows) { r -= p.rows; b++; }
                    const int64_t zrows = (p.rows + 7) >> 3;
                    atomicOr((unsigned int*)&p.zp_packed[(b * zrows + (r >> 3)) * p.groups_per_row + k],
                             ((uint32_t)((int)z + 8) & 0xfu) << (4 * (int)(r & 7)));
                }
            }
            // ---- C. quantize + pack into the staging buffer
            if (gi == 0) {
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
            }
            const bool unsafe = (QT == QT_FP4) ? !fp4_scale_is_safe(s) : !scale_is_safe(__float_as_uint(s));
            const bool add_zp = (QT == QT_FP8) ? (p.has_zp != 0) : true;
            const uint32_t z2 = (__float_as_uint(z) >> 16) * 0x10001u;
            const uint32_t oaddr = out_base + (uint32_t)gl * (N * OUT_CHUNK) + (uint32_t)half * (NL * OUT_CHUNK);
            uint32_t diff = 0;
            // Measured (scripts/ab_tma.py, fraction of the HBM roofline; asym g128 / sym g128 / sym g32 / fp8 g32):
            //   bracket check by HFMA2, repair in the loop     0.67 / 0.84 / 0.79 / 0.95
            //   bracket check by XOR,   repair in the loop       -  / 0.74 / 0.84 / 0.95
            //   bracket check by XOR,   repair deferred        0.72 / 0.74 / 0.80 / 0.95
            //   bracket check by HFMA2, repair deferred        0.69 / 0.77 / 0.83 / 0.95
            // -> per-format choice below (overridable at build time: B200Q_NVCC_DEFS="-DB200Q_TMA_DEFER=... -DB200Q_TMA_HDIFF=...").
            // DEFER (asymmetric INT4, the longest per-element chain): the exact repair runs once per group after the loop, which
            // keeps the loop call-free straight-line code (no per-chunk branch / BSSY / constant re-materialisation): +9 % measured.
            // The other formats are faster with the repair inside the loop (a group-wide flag fires for ~0.2 % of the groups, i.e.
            // for ~6 % of the warp iterations, and the divergent repair pass then costs more than the branches saved).
            // Batches of QB chunks: enough independent chains to hide the conversion latencies without letting the compiler hoist
            // all 16 loads of a group (which drove the kernel to the 168-register cap).
#ifndef B200Q_TMA_DEFER
#define B200Q_TMA_DEFER (QT == QT_INT && !SYM)
#endif
#ifndef B200Q_TMA_HDIFF
#define B200Q_TMA_HDIFF (QT == QT_INT && SYM && LOG2N == 4)
#endif
            constexpr bool DEFER = ONE || (B200Q_TMA_DEFER);
            constexpr bool HDIFF = FMA && QT != QT_FP4 && (B200Q_TMA_HDIFF);
            constexpr int QB = !DEFER ? NL : (NL < 4 ? NL : 4);
            // chunk i of the lane's walk sits at position c = i ^ rot.  The stage and group bases are multiples of NL * 16 bytes, so
            // base + (c << 4) == (base + (rot << 4)) ^ (i << 4), and with i = i0 + j (i0 a multiple of QB, j < QB compile time) the
            // per-chunk address is ONE LOP3 with an immediate (the rotate-and-mask form cost 3 instructions per address).
            const uint32_t gswz = gaddr + ((uint32_t)rot << 4), oswz = oaddr + (uint32_t)rot * OUT_CHUNK;
#pragma unroll 1
            for (int i0 = 0; i0 < NL; i0 += QB) {
            const uint32_t gblk = gswz ^ ((uint32_t)i0 << 4), oblk = oswz ^ ((uint32_t)i0 * OUT_CHUNK);
            uint4 vq[QB];
            if (ONE && PF) {  // the batch's loads first: one exposed LDS latency per QB chunks
#pragma unroll
                for (int j = 0; j < QB; j++) vq[j] = lds128(gblk ^ ((uint32_t)j << 4));
            }
#pragma unroll
            for (int i = i0; i < i0 + QB; i++) {
                const int c = i ^ rot;
                const uint4 v = (ONE && PF) ? vq[i - i0] : lds128(gblk ^ ((uint32_t)(i - i0) << 4));
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                uint32_t h[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const f32x2 x = FMA ? bf16x2_to_f32x2_fma(w[k]) : bf16x2_to_f32x2(w[k]);
                    float al, ah, bl, bh;
                    if (ONE && QT == QT_INT) {
                        unpack2(mul2(x, br.lo), al, ah);
                        uint32_t u = cvt_bf16x2(ah, al);
                        if (!SYM) u = hadd2(u, z2);
                        h[k] = __viaddmin_s16x2_relu(hadd2(u, kMagic), kUnbiasR, 0x000f000fu);
                    } else if (ONE && QT == QT_FP8) {
                        unpack2(add_zp ? mul2_plus0(x, br.lo) : mul2(x, br.lo), al, ah);
                        const uint32_t u = cvt_bf16x2(ah, al);
                        if (FMA) {
                            float ul, uh;
                            unpack2(bf16x2_to_f32x2_fma(u), ul, uh);
                            h[k] = cvt_e4m3x2(uh, ul);
                        } else {
                            h[k] = cvt_e4m3x2(__uint_as_float(u & 0xffff0000u), __uint_as_float(u << 16));
                        }
                    } else if (QT == QT_INT) {
                        unpack2(mul2(x, br.lo), al, ah);
                        unpack2(mul2(x, br.hi), bl, bh);
                        uint32_t u = cvt_bf16x2(ah, al);
                        if (HDIFF) diff = hdiff2_acc(u, cvt_bf16x2(bh, bl), diff);
                        else diff |= u ^ cvt_bf16x2(bh, bl);
                        if (!SYM) u = hadd2(u, z2);
                        h[k] = __viaddmin_s16x2_relu(hadd2(u, kMagic), kUnbias, 0x000f000fu);
                    } else if (QT == QT_FP8) {
                        unpack2(add_zp ? mul2_plus0(x, br.lo) : mul2(x, br.lo), al, ah);
                        unpack2(add_zp ? mul2_plus0(x, br.hi) : mul2(x, br.hi), bl, bh);
                        const uint32_t u = cvt_bf16x2(ah, al);
                        if (HDIFF) diff = hdiff2_acc(u, cvt_bf16x2(bh, bl), diff);
                        else diff |= u ^ cvt_bf16x2(bh, bl);
                        if (FMA) {
                            float ul, uh;
                            unpack2(bf16x2_to_f32x2_fma(u), ul, uh);
                            h[k] = cvt_e4m3x2(uh, ul);
                        } else {
                            h[k] = cvt_e4m3x2(__uint_as_float(u & 0xffff0000u), __uint_as_float(u << 16));
                        }
                    } else {
                        unpack2(mul2_plus0(x, br.lo), al, ah);
                        unpack2(mul2_plus0(x, br.hi), bl, bh);
                        h[k] = cvt_e2m1x2(ah, al);
                        diff |= h[k] ^ cvt_e2m1x2(bh, bl);
                    }
                }
                uint2 packed;
                if (QT == QT_INT) {
                    const uint32_t x01 = prmt(h[0], h[1], 0x6420), x23 = prmt(h[2], h[3], 0x6420);
                    if (FMA) packed = make_uint2(prmt(fold_nibbles(x01), fold_nibbles(x23), 0x6420), 0u);
                    else packed = make_uint2(prmt(x01 | (x01 >> 4), x23 | (x23 >> 4), 0x6420), 0u);
                } else if (QT == QT_FP8) {
                    packed = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
                } else {
                    packed = make_uint2(h[0] | (h[1] << 8) | (h[2] << 16) | (h[3] << 24), 0u);
                }
                if (!DEFER) {
                    if ((HDIFF ? hdiff2_any(diff) : diff != 0) || unsafe) packed = repair_chunk<QT, SYM>(v, s, z, add_zp, unsafe, packed);
                    diff = 0;
                }
                (void)c;
                if (QT == QT_FP8) sts64(oblk ^ ((uint32_t)(i - i0) * 8), packed);
                else sts32(oblk ^ ((uint32_t)(i - i0) * 4), packed.x);
            }
            }
            // ---- C'. exact repair, once per group and out of the hot loop (rare: p ~ 1e-4 per element).  `diff` collected every
            // disagreement of the two bracket ends over the lane's chunks; the loop above is call-free straight-line code, so the
            // compiler interleaves chunks and keeps its constants in registers.  The inputs are still in the stage.
            if (DEFER && (ONE ? unsafe : ((HDIFF ? hdiff2_any(diff) : diff != 0) || unsafe))) {
#pragma unroll 1
                for (int c = 0; c < NL; c++) {
                    const uint4 v = lds128(gaddr + (c << 4));
                    uint2 packed;
                    if (QT == QT_FP8) packed = lds64(oaddr + c * 8);
                    else packed = make_uint2(lds32(oaddr + c * 4), 0u);
                    packed = repair_chunk<QT, SYM>(v, s, z, add_zp, unsafe, packed);
                    if (QT == QT_FP8) sts64(oaddr + c * 8, packed);
                    else sts32(oaddr + c * 4, packed.x);
                }
            }
        }
        // ---- D. hand the packed tile to the TMA engine, refill this input stage
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_1d(p.out + g0 * (N * OUT_CHUNK), out_base, (uint32_t)n_here * (N * OUT_CHUNK));
            bulk_commit();
            const int64_t nt = tile + (int64_t)kStages * wstride;
            if (nt < n_tiles) issue(nt, stage);
        }
    }
    if (lane == 0) bulk_wait0();
}

// asym INT4: int8 zero points [batch, rows, gpr] -> CT's row-packed layout int32 [batch, ceil(rows / 8), gpr] (pack_to_int32 with
// packed_dim = 0, CT:compressors/pack_quantized/base.py:70-73).  One thread per output word; the 8 byte loads of a warp are 32
// consecutive bytes each.  Replaces one atomicOr per group + a memset of the packed buffer (round-1 verdict, weak #2c).
__global__ void zp_pack_rows_kernel(const int8_t* __restrict__ zp, int64_t batch, int64_t rows, int64_t gpr, int32_t* __restrict__ out) {
    const int64_t zrows = (rows + 7) >> 3;
    if ((gpr & 3) == 0 && (((uintptr_t)zp) & 3) == 0 && (((uintptr_t)out) & 15) == 0) {
        // four group columns per thread: 8 aligned 4-byte loads (one per row of the octet), SWAR nibble placement, one 16-byte store
        const int64_t q = gpr >> 2, n = batch * zrows * q;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t k4 = i % q, t = i / q, zr = t % zrows, b = t / zrows;
            const int8_t* src = zp + (b * rows + zr * 8) * gpr + k4 * 4;
            const int nr = (int)min((int64_t)8, rows - zr * 8);
            uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j < nr) {
                    // (zp + 8) & 0xf per byte without carries between the bytes: zp in [-8, 7], so its low nibble XOR 8 is zp + 8
                    const uint32_t v = ((*reinterpret_cast<const uint32_t*>(src + (int64_t)j * gpr)) & 0x0f0f0f0fu) ^ 0x08080808u;
                    w0 |= (v & 0xffu) << (4 * j);
                    w1 |= ((v >> 8) & 0xffu) << (4 * j);
                    w2 |= ((v >> 16) & 0xffu) << (4 * j);
                    w3 |= (v >> 24) << (4 * j);
                }
            }
            *reinterpret_cast<uint4*>(out + (b * zrows + zr) * gpr + k4 * 4) = make_uint4(w0, w1, w2, w3);
        }
        return;
    }
    const int64_t n = batch * zrows * gpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i % gpr, t = i / gpr, zr = t % zrows, b = t / zrows;
        const int8_t* src = zp + (b * rows + zr * 8) * gpr + k;
        const int nr = (int)min((int64_t)8, rows - zr * 8);
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j < nr) word |= ((uint32_t)((int)src[(int64_t)j * gpr] + 8) & 0xfu) << (4 * j);
        out[i] = (int32_t)word;
    }
}

template <int QT, bool SYM, int LOG2N, bool FMA, int TILE, bool ONE, bool PF = true, int STAGES = kStages, bool SPIN = false, int NWARPS = 0>
int launch_tma_v(const TmaParams& p, cudaStream_t st) {
    constexpr int OUT_BYTES = (TILE / 8) * ((QT == QT_FP8) ? 8 : 4);
    constexpr int kWarps = WarpsFor<QT, TILE, STAGES, NWARPS>::value;
    const size_t smem = (size_t)kWarps * (STAGES * TILE * 2 + OUT_BYTES + 256);
    static bool configured = false;  // benign race: idempotent attribute
    if (!configured) {
        cudaFuncSetAttribute(group_tma_kernel<QT, SYM, LOG2N, FMA, TILE, false, ONE, PF, STAGES, SPIN, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    constexpr int GPT = TILE / (8 << LOG2N);
    const int64_t n_tiles = (p.n_groups + GPT - 1) / GPT;
    const int64_t ctas = max((int64_t)1, min((int64_t)kNumSMs, (n_tiles + kWarps - 1) / kWarps));
    group_tma_kernel<QT, SYM, LOG2N, FMA, TILE, false, ONE, PF, STAGES, SPIN, NWARPS><<<(unsigned)ctas, kWarps * 32, smem, st>>>(p);
    B200Q_CHECK_LAUNCH();
    if (QT == QT_INT && !SYM && p.zp_i8 != nullptr) {
        const int64_t words = (p.n_groups / p.groups_per_mat) * ((p.rows + 7) >> 3) * p.groups_per_row;
        const int64_t blocks = min((int64_t)kNumSMs * 8, (words / ((p.groups_per_row & 3) == 0 ? 4 : 1) + 255) / 256);
        zp_pack_rows_kernel<<<(unsigned)max((int64_t)1, blocks), 256, 0, st>>>(p.zp_i8, p.n_groups / p.groups_per_mat, p.rows, p.groups_per_row,
                                                                                p.zp_packed);
        B200Q_CHECK_LAUNCH();
    }
    return B200Q_OK;
}

// Round 1 (bracketed reciprocal, both ends through cvt; scripts/ab_tma.py, 16 x [9728, 2560] bf16, fraction of 6.55 TB/s):
//                        8 KB tiles, 12 warps          4 KB tiles, 24 warps
//   INT4 g128 asym       0.668 -> 0.680 (FMA mix)      0.680
//   INT4 g128 sym        0.819 -> 0.836                0.748
//   INT4 g32  sym        0.776 -> 0.793                0.781
//   FP8  g32             0.950 -> 0.913                0.839
// Round 2: INT4 / FP8 take the single-evaluation variant (ONE) by default.  B200Q_TMA_BRACKET=1 selects the round-1 bracket kernels,
// B200Q_TMA_TILE=2048, B200Q_TMA_LEGACY_ALU=1 / B200Q_TMA_FMA=1 the other variants (read once per process; identical bits).
template <int QT, bool SYM, int LOG2N>
int launch_tma(const TmaParams& p, cudaStream_t st) {
    static const int tile = getenv("B200Q_TMA_TILE") ? atoi(getenv("B200Q_TMA_TILE")) : 4096;
    static const bool fma = getenv("B200Q_TMA_LEGACY_ALU") ? false : (getenv("B200Q_TMA_FMA") ? true : QT == QT_INT);
    static const bool bracket = getenv("B200Q_TMA_BRACKET") != nullptr;
    // A/B variants of the INT4 kernels (identical bits): B200Q_TMA_VAR = 1 ONE without the batched loads, 2 / 3 ONE with a 3- / 4-deep
    // ring and 8 / 6 warps, 4 / 5 the bracket kernel with a 3- / 4-deep ring
    static const int var = getenv("B200Q_TMA_VAR") ? atoi(getenv("B200Q_TMA_VAR")) : 0;
    if constexpr (QT == QT_INT) {
        switch (var) {
        case 1: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, false, 2>(p, st);
        case 2: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 3>(p, st);
        case 3: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 4>(p, st);
        case 4: return launch_tma_v<QT, SYM, LOG2N, true, 4096, false, true, 3>(p, st);
        case 5: return launch_tma_v<QT, SYM, LOG2N, true, 4096, false, true, 4>(p, st);
        case 6: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, true>(p, st);    // ONE, test_wait spin
        case 7: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 3, true>(p, st);    // ONE, 3 stages, spin
        // fewer warps = fewer 8 KB loads in flight (ncu: ONE sits on the mbarrier 1.7 warps / issue vs 0.09 for the bracket kernel,
        // yet gets 9 % LESS HBM bandwidth -- 28 MB of outstanding bulk loads may be past the DRAM scheduler's sweet spot)
        case 8: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, false, 8>(p, st);
        case 9: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, false, 6>(p, st);
        case 10: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, false, 4>(p, st);
        default: break;
        }
    }
    if constexpr (QT == QT_INT) {
        // Default (round 2), INT4: the single-evaluation kernel with EIGHT warps per SM.  With 12 warps x 2 stages x 8 KB the lean kernel keeps
        // ~28 MB of bulk loads outstanding and gets 9 % LESS bandwidth than the heavier bracket kernel (ncu: 1.7 warps per issue parked
        // on the mbarrier vs 0.09); 8 warps: 6.7-7.0 TB/s on every INT4 scheme, 6 warps: compute-bound again (scripts/ab_tma2.py,
        // profiles/r2_int4_kernel_ab.md).  The one compute-heavy INT4 variant, asymmetric g32 (a qparam chain + zero point per 32
        // elements), keeps 12 warps: 0.88 of the roofline vs 0.73 with 8.  FP8 GROUP stays on the bracket kernel with 10 warps
        // (g32: 0.93 vs 0.83-0.87 for the single-evaluation variants -- its bf16 -> e4m3 second conversion needs the issue slots).
        // B200Q_TMA_WARPS = 7 / 9 / 10 / 12 for sweeps.
        static const int warps = getenv("B200Q_TMA_WARPS") ? atoi(getenv("B200Q_TMA_WARPS")) : ((LOG2N <= 2 && !SYM) ? 12 : 8);
        if (!bracket) {
            if (fma) {
                switch (warps) {
                case 7: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, false, 7>(p, st);
                case 9: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, false, 9>(p, st);
                case 10: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, false, 10>(p, st);
                case 12: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, false, 12>(p, st);
                default: return launch_tma_v<QT, SYM, LOG2N, true, 4096, true, true, 2, false, 8>(p, st);
                }
            }
            switch (warps) {
            case 7: return launch_tma_v<QT, SYM, LOG2N, false, 4096, true, true, 2, false, 7>(p, st);
            case 9: return launch_tma_v<QT, SYM, LOG2N, false, 4096, true, true, 2, false, 9>(p, st);
            case 10: return launch_tma_v<QT, SYM, LOG2N, false, 4096, true, true, 2, false, 10>(p, st);
            case 12: return launch_tma_v<QT, SYM, LOG2N, false, 4096, true, true, 2, false, 12>(p, st);
            default: return launch_tma_v<QT, SYM, LOG2N, false, 4096, true, true, 2, false, 8>(p, st);
            }
        }
        if (tile == 2048) return launch_tma_v<QT, SYM, LOG2N, true, 2048, false>(p, st);
    }
    return fma ? launch_tma_v<QT, SYM, LOG2N, true, 4096, false>(p, st) : launch_tma_v<QT, SYM, LOG2N, false, 4096, false>(p, st);
}

template <int QT, bool SYM, int LOG2N>
int launch_tma_supplied(const TmaParams& p, cudaStream_t st) {
    constexpr int TILE = 4096;
    constexpr bool FMA = QT == QT_INT;
    constexpr bool ONE = QT != QT_FP4;
    constexpr int OUT_BYTES = (TILE / 8) * ((QT == QT_FP8) ? 8 : 4);
    constexpr int kWarps = WarpsFor<QT, TILE>::value;
    const size_t smem = (size_t)kWarps * (kStages * TILE * 2 + OUT_BYTES + 256);
    static bool configured = false;  // benign race: idempotent attribute
    if (!configured) {
        cudaFuncSetAttribute(group_tma_kernel<QT, SYM, LOG2N, FMA, TILE, true, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    constexpr int GPT = TILE / (8 << LOG2N);
    const int64_t n_tiles = (p.n_groups + GPT - 1) / GPT;
    const int64_t ctas = max((int64_t)1, min((int64_t)kNumSMs, (n_tiles + kWarps - 1) / kWarps));
    group_tma_kernel<QT, SYM, LOG2N, FMA, TILE, true, ONE><<<(unsigned)ctas, kWarps * 32, smem, st>>>(p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace

// bf16 only, caller-supplied qparams (MODE_QUANT_PACK): scale bf16 [n_groups], zp int8 [n_groups] or null.  INT4 and FP8 GROUP.
int launch_group_tma_supplied(int qt, const GroupParams& gp, int64_t batch, cudaStream_t st) {
    const int g = gp.group;
    if (!(g == 32 || g == 64 || g == 128) || gp.cols % g != 0) return B200Q_ENOSYS;
    if ((((uintptr_t)gp.w) & 15) != 0 || (((uintptr_t)gp.out) & 15) != 0 || batch * gp.rows * gp.cols == 0) return B200Q_ENOSYS;
    if ((batch * gp.rows * gp.cols) % 32 != 0 || (((uintptr_t)gp.scale) & 1) != 0) return B200Q_ENOSYS;
    if (gp.cols / g >= (1ll << 30) || gp.rows >= (1ll << 40)) return B200Q_ENOSYS;
    TmaParams p{};
    p.w = (const uint16_t*)gp.w;
    p.groups_per_row = gp.cols / g;
    p.groups_per_mat = gp.rows * p.groups_per_row;
    p.n_groups = batch * p.groups_per_mat;
    p.rows = gp.rows;
    p.out = (uint8_t*)gp.out;
    p.scale = gp.scale;
    p.zp_in = gp.zp_in;
    p.has_zp = gp.has_zp;
    if (qt == QT_INT && gp.nbits == 4) {
        const bool sym = gp.zp_in == nullptr;  // quant_int adds the zero point only when one is supplied
        switch (g) {
        case 32: return sym ? launch_tma_supplied<QT_INT, true, 2>(p, st) : launch_tma_supplied<QT_INT, false, 2>(p, st);
        case 64: return sym ? launch_tma_supplied<QT_INT, true, 3>(p, st) : launch_tma_supplied<QT_INT, false, 3>(p, st);
        default: return sym ? launch_tma_supplied<QT_INT, true, 4>(p, st) : launch_tma_supplied<QT_INT, false, 4>(p, st);
        }
    }
    if (qt == QT_FP8) {
        switch (g) {
        case 32: return launch_tma_supplied<QT_FP8, true, 2>(p, st);
        case 64: return launch_tma_supplied<QT_FP8, true, 3>(p, st);
        default: return launch_tma_supplied<QT_FP8, true, 4>(p, st);
        }
    }
    return B200Q_ENOSYS;
}

// bf16 only.  Returns B200Q_ENOSYS when the scheme / shape is not covered.
int launch_group_tma(int qt, const GroupParams& gp, int64_t batch, cudaStream_t st) {
    const int g = gp.group;
    if (!(g == 16 || g == 32 || g == 64 || g == 128) || gp.cols % g != 0) return B200Q_ENOSYS;
    if ((((uintptr_t)gp.w) & 15) != 0 || (((uintptr_t)gp.out) & 15) != 0 || batch * gp.rows * gp.cols == 0) return B200Q_ENOSYS;
    if ((batch * gp.rows * gp.cols) % 32 != 0) return B200Q_ENOSYS;  // bulk copies move multiples of 16 bytes
    if (gp.cols / g >= (1ll << 30) || gp.rows >= (1ll << 40)) return B200Q_ENOSYS;
    TmaParams p{};
    p.w = (const uint16_t*)gp.w;
    p.groups_per_row = gp.cols / g;
    p.groups_per_mat = gp.rows * p.groups_per_row;
    p.n_groups = batch * p.groups_per_mat;
    p.rows = gp.rows;
    p.out = (uint8_t*)gp.out;
    p.scale = gp.scale;
    p.zp_packed = gp.zp_packed;
    p.zp_i8 = gp.zp_scratch;
    p.gs = gp.gs;
    p.gs_stride = gp.gs_stride;
    p.has_zp = gp.has_zp;
    if (qt == QT_INT && gp.nbits == 4) {
        if (!gp.symmetric) {
            if (gp.zp_packed == nullptr) return B200Q_ENOSYS;
            if (gp.zp_scratch == nullptr)  // legacy: atomics into a zeroed buffer
                cudaMemsetAsync(gp.zp_packed, 0, sizeof(int32_t) * batch * ((gp.rows + 7) / 8) * p.groups_per_row, st);
        }
        switch (g) {
        case 32: return gp.symmetric ? launch_tma<QT_INT, true, 2>(p, st) : launch_tma<QT_INT, false, 2>(p, st);
        case 64: return gp.symmetric ? launch_tma<QT_INT, true, 3>(p, st) : launch_tma<QT_INT, false, 3>(p, st);
        case 128: return gp.symmetric ? launch_tma<QT_INT, true, 4>(p, st) : launch_tma<QT_INT, false, 4>(p, st);
        default: return B200Q_ENOSYS;
        }
    }
    if (qt == QT_FP8) {
        switch (g) {
        case 32: return launch_tma<QT_FP8, true, 2>(p, st);
        case 64: return launch_tma<QT_FP8, true, 3>(p, st);
        case 128: return launch_tma<QT_FP8, true, 4>(p, st);
        default: return B200Q_ENOSYS;
        }
    }
    // NVFP4 (g16): measured slower than the thread-owns-group register kernel (quant_tile_fast.cu: 122 us vs 173 us on
    // 128 experts) -- the per-group chain dominates at 16 elements per group; opt-in for experiments only.
    if (qt == QT_FP4 && g == 16 && getenv("B200Q_FP4_TMA") != nullptr) return launch_tma<QT_FP4, true, 1>(p, st);
    return B200Q_ENOSYS;
}

}  // namespace b200q
