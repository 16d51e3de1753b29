// decode_fast.cu -- bf16 decode (compressed tensors -> weights) at HBM speed: SURVEY.md §8f rank 1.
//
//   INT4 pack-quantized   weight_packed int32 + weight_scale bf16 (+ row-packed weight_zero_point)   CT:compressors/pack_quantized/base.py:79-113
//   FP8  float-quantized  weight e4m3 + weight_scale bf16 (channel / group / block / tensor)         CT:quantization/lifecycle/forward.py:77-145
//   NVFP4                 weight_packed u8 + weight_scale e4m3 + weight_global_scale fp32            CT:compressors/nvfp4/base.py:74-96
//
// The generic kernels (pack.cu, quant_elementwise.cu) index every element with 64-bit divisions and run-time bit widths; they
// measured 0.06-0.17 of the HBM roofline.  Here a warp walks one row (INT4 / FP8) or a flat run of groups (NVFP4): everything
// that depends on the row is computed once per row, loads are 4- or 8-byte words contiguous across the lanes, stores are 16-byte
// vectors contiguous across the lanes, and the arithmetic is packed:
//   INT4: the nibble n becomes the bf16 128 + n by OR-ing it into 0x4300 (one ulp == 1 in [128, 256)); subtracting the bf16
//         136 + zp gives (n - 8) - zp exactly, and mul.rn.bf16x2 rounds the product with the scale once -- the reference's
//         T(q) - T(zp), then * scale rounded to T (forward_helpers.py:258-266);
//   FP8 / FP4: cvt.rn.f16x2.{e4m3x2,e2m1x2} is exact, the fp32 product with the scale is exact (or rounded once for NVFP4's fp32
//         scale), cvt.rn.bf16x2.f32 is the final cast to T.
#include <cstdlib>
#include "../../include/b200q.h"
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.cuh"

namespace b200q {
namespace {
using namespace fast;

__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t e4m3x2_to_f16x2(uint32_t two_bytes) {
    uint32_t r;
    asm("{ .reg .b16 t; cvt.u16.u32 t, %1; cvt.rn.f16x2.e4m3x2 %0, t; }" : "=r"(r) : "r"(two_bytes));
    return r;
}
__device__ __forceinline__ uint32_t e2m1x2_to_f16x2(uint32_t one_byte) {
    uint32_t r;
    asm("{ .reg .b8 t; .reg .b16 u; cvt.u16.u32 u, %1; cvt.u8.u16 t, u; cvt.rn.f16x2.e2m1x2 %0, t; }" : "=r"(r) : "r"(one_byte));
    return r;
}
// f16x2 (exact small-format values) * fp32 scale -> bf16x2, one rounding per product and one for the cast
__device__ __forceinline__ uint32_t f16x2_scale_to_bf16x2(uint32_t h2, float s) {
    float lo, hi;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(lo), "=f"(hi) : "r"(h2));
    return cvt_bf16x2(__fmul_rn(hi, s), __fmul_rn(lo, s));
}

// ------------------------------------------------------------------------------------------------ INT4
// one warp per row; lane handles packed words lane, lane + 32, ... (8 elements each: 4-byte load, 16-byte store)
template <bool HAS_ZP>
__global__ void __launch_bounds__(256) decode_int4_kernel(const uint32_t* __restrict__ packed, const uint16_t* __restrict__ scale,
                                                          const uint32_t* __restrict__ zp_packed, int64_t total_rows, int64_t rows, int cols,
                                                          int wpg /* words per group, 0 = channel */, uint4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const int wpr = cols >> 3, gtot = wpg ? wpr / wpg : 1;
    const int sh = (wpg && (wpg & (wpg - 1)) == 0) ? __ffs(wpg) - 1 : -1;
    // (matrix, row) of this warp's current row, advanced without a 64-bit division per row
    const int64_t step_b = nwarps / rows, step_r = nwarps - step_b * rows;
    int64_t b = HAS_ZP ? warp / rows : 0, r = HAS_ZP ? warp - b * rows : 0;
    for (int64_t br = warp; br < total_rows; br += nwarps) {
        const uint32_t* prow = packed + br * wpr;
        const uint16_t* srow = scale + br * gtot;
        const uint32_t* zrow = nullptr;
        int zshift = 0;
        if (HAS_ZP) {
            zrow = zp_packed + (b * ((rows + 7) >> 3) + (r >> 3)) * gtot;
            zshift = 4 * (int)(r & 7);
            b += step_b;
            r += step_r;
            if (r >= rows) { r -= rows; b++; }
        }
        uint4* orow = out + br * wpr;
#pragma unroll 4
        for (int cw = lane; cw < wpr; cw += 32) {
            const uint32_t w = __ldg(prow + cw);
            const int g = wpg ? (sh >= 0 ? cw >> sh : cw / wpg) : 0;
            const uint32_t s2 = (uint32_t)__ldg(srow + g) * 0x10001u;
            uint32_t c2 = 0x43084308u;  // bf16x2 (136, 136): the offset-binary bias on top of the 128 of the nibble trick
            if (HAS_ZP) c2 = (0x4300u | ((__ldg(zrow + g) >> zshift) & 0xfu)) * 0x10001u;  // 128 + (zp + 8)
            uint32_t y[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t t = w >> (8 * k);
                const uint32_t pair = 0x43004300u | (t & 0xfu) | ((t & 0xf0u) << 12);
                y[k] = hmul2(hsub2(pair, c2), s2);
            }
            stg_stream(orow + cw, make_uint4(y[0], y[1], y[2], y[3]));
        }
    }
}

// Power-of-two group sizes (every recipe: 32, 128): the same row walk in BATCHES of U words per lane.  The kernel above leaves the
// `shift or divide` choice and the dependent scale / zero-point loads inside the loop, and ptxas then serialises word -> scale ->
// convert -> store with one word of prefetch: each warp has ~2 loads in flight and the kernel is latency-bound (the 48-register
// asymmetric variant, 5 CTAs per SM, measured 0.66-0.73 of the HBM roofline against 0.86 symmetric).  Here all 2U (3U) loads of a
// batch are independent of each other and issued before the first conversion.
template <bool HAS_ZP, int U>
__global__ void __launch_bounds__(256) decode_int4_batch_kernel(const uint32_t* __restrict__ packed, const uint16_t* __restrict__ scale,
                                                                const uint32_t* __restrict__ zp_packed, int64_t total_rows, int64_t rows, int cols,
                                                                int sh /* log2(words per group) */, uint4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const int wpr = cols >> 3, gtot = wpr >> sh;
    const int64_t step_b = nwarps / rows, step_r = nwarps - step_b * rows;
    int64_t b = HAS_ZP ? warp / rows : 0, r = HAS_ZP ? warp - b * rows : 0;
    for (int64_t br = warp; br < total_rows; br += nwarps) {
        const uint32_t* prow = packed + br * wpr;
        const uint16_t* srow = scale + br * gtot;
        const uint32_t* zrow = nullptr;
        int zshift = 0;
        if (HAS_ZP) {
            zrow = zp_packed + (b * ((rows + 7) >> 3) + (r >> 3)) * gtot;
            zshift = 4 * (int)(r & 7);
            b += step_b;
            r += step_r;
            if (r >= rows) { r -= rows; b++; }
        }
        uint4* orow = out + br * wpr;
        for (int c0 = lane; c0 < wpr; c0 += 32 * U) {
            uint32_t w[U], sc[U], zw[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int cw = c0 + 32 * u;
                const bool ok = cw < wpr;
                const int g = cw >> sh;
                w[u] = ok ? __ldg(prow + cw) : 0u;
                sc[u] = ok ? (uint32_t)__ldg(srow + g) : 0u;
                zw[u] = (HAS_ZP && ok) ? __ldg(zrow + g) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int cw = c0 + 32 * u;
                const uint32_t s2 = sc[u] * 0x10001u;
                uint32_t c2 = 0x43084308u;  // bf16x2 (136, 136)
                if (HAS_ZP) c2 = (0x4300u | ((zw[u] >> zshift) & 0xfu)) * 0x10001u;  // 128 + (zp + 8)
                uint32_t y[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t t = w[u] >> (8 * k);
                    const uint32_t pair = 0x43004300u | (t & 0xfu) | ((t & 0xf0u) << 12);
                    y[k] = hmul2(hsub2(pair, c2), s2);
                }
                if (cw < wpr) stg_stream(orow + cw, make_uint4(y[0], y[1], y[2], y[3]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ FP8
// one warp per row; lane handles 8-code chunks lane, lane + 32, ... (8-byte load, 16-byte store).  Scale index of chunk cw:
//   s_row_stride * (row / rows_per_scale) + cw / chunks_per_scale
__global__ void __launch_bounds__(256) decode_fp8_kernel(const uint2* __restrict__ codes, const uint16_t* __restrict__ scale, int64_t rows, int cols,
                                                         int rows_per_scale, int chunks_per_scale, int64_t s_row_stride, uint4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const int cpr = cols >> 3;
    const int sh = (chunks_per_scale & (chunks_per_scale - 1)) == 0 ? __ffs(chunks_per_scale) - 1 : -1;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const uint2* crow = codes + r * cpr;
        const uint16_t* srow = scale + (r / rows_per_scale) * s_row_stride;
        uint4* orow = out + r * cpr;
#pragma unroll 4
        for (int cw = lane; cw < cpr; cw += 32) {
            const uint2 c = __ldg(crow + cw);
            const int si = sh >= 0 ? cw >> sh : cw / chunks_per_scale;
            const float s = __uint_as_float((uint32_t)__ldg(srow + si) << 16);
            uint4 y;
            y.x = f16x2_scale_to_bf16x2(e4m3x2_to_f16x2(c.x & 0xffffu), s);
            y.y = f16x2_scale_to_bf16x2(e4m3x2_to_f16x2(c.x >> 16), s);
            y.z = f16x2_scale_to_bf16x2(e4m3x2_to_f16x2(c.y & 0xffffu), s);
            y.w = f16x2_scale_to_bf16x2(e4m3x2_to_f16x2(c.y >> 16), s);
            stg_stream(orow + cw, y);
        }
    }
}

// Batched variant (see decode_int4_batch_kernel): U chunks per lane, every load of a batch issued before the first conversion.
// ROW_SCALE: one scale for the whole row (channel / tensor / block or group at least as wide as the row) -> loaded once per row;
// otherwise chunks_per_scale is a power of two (sh).
template <bool ROW_SCALE, int U>
__global__ void __launch_bounds__(256) decode_fp8_batch_kernel(const uint2* __restrict__ codes, const uint16_t* __restrict__ scale, int64_t rows, int cols,
                                                               int rows_per_scale, int sh, int64_t s_row_stride, uint4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const int cpr = cols >> 3;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const uint2* crow = codes + r * cpr;
        const uint16_t* srow = scale + (r / rows_per_scale) * s_row_stride;
        uint4* orow = out + r * cpr;
        uint32_t s_row = 0;
        if (ROW_SCALE) s_row = (uint32_t)__ldg(srow) << 16;
        for (int c0 = lane; c0 < cpr; c0 += 32 * U) {
            uint2 c[U];
            uint32_t sc[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int cw = c0 + 32 * u;
                const bool ok = cw < cpr;
                c[u] = ok ? __ldg(crow + cw) : make_uint2(0u, 0u);
                sc[u] = ROW_SCALE ? s_row : (ok ? (uint32_t)__ldg(srow + (cw >> sh)) << 16 : 0u);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int cw = c0 + 32 * u;
                const float s = __uint_as_float(sc[u]);
                uint4 y;
                y.x = f16x2_scale_to_bf16x2(e4m3x2_to_f16x2(c[u].x & 0xffffu), s);
                y.y = f16x2_scale_to_bf16x2(e4m3x2_to_f16x2(c[u].x >> 16), s);
                y.z = f16x2_scale_to_bf16x2(e4m3x2_to_f16x2(c[u].y & 0xffffu), s);
                y.w = f16x2_scale_to_bf16x2(e4m3x2_to_f16x2(c[u].y >> 16), s);
                if (cw < cpr) stg_stream(orow + cw, y);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ NVFP4
// blockIdx.y = matrix (one global scale): table e4m3 code -> fp32(code) / gs built once per CTA; thread handles one 32-bit word
// (8 codes, half a group): 4-byte load, 16-byte store
__global__ void __launch_bounds__(256) decode_nvfp4_kernel(const uint32_t* __restrict__ packed, const uint8_t* __restrict__ scale,
                                                           const float* __restrict__ gs, int gs_stride, int64_t words_per_mat, uint4* __restrict__ out) {
    __shared__ float s_eff[256];
    const int64_t b = blockIdx.y;
    s_eff[threadIdx.x] = fdiv(e4m3_decode((uint8_t)threadIdx.x), gs[gs_stride ? b : 0]);
    __syncthreads();
    const uint32_t* pm = packed + b * words_per_mat;
    const uint8_t* sm = scale + b * (words_per_mat >> 1);
    uint4* om = out + b * words_per_mat;
    constexpr int U = 4;  // words per thread per tile: all loads of a tile are issued before the first conversion
    const int64_t n_tiles = (words_per_mat + 256 * U - 1) / (256 * U);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t w0 = tile * (256 * U) + threadIdx.x;
        uint32_t c[U];
        uint32_t sc[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t w = w0 + u * 256;
            const bool ok = w < words_per_mat;
            c[u] = ok ? __ldg(pm + w) : 0u;
            sc[u] = ok ? (uint32_t)__ldg(sm + (w >> 1)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t w = w0 + u * 256;
            if (w >= words_per_mat) continue;
            const float s = s_eff[sc[u]];
            uint4 y;
            y.x = f16x2_scale_to_bf16x2(e2m1x2_to_f16x2(c[u] & 0xffu), s);
            y.y = f16x2_scale_to_bf16x2(e2m1x2_to_f16x2((c[u] >> 8) & 0xffu), s);
            y.z = f16x2_scale_to_bf16x2(e2m1x2_to_f16x2((c[u] >> 16) & 0xffu), s);
            y.w = f16x2_scale_to_bf16x2(e2m1x2_to_f16x2(c[u] >> 24), s);
            stg_stream(om + w, y);
        }
    }
}

// CTAs per SM a decode grid is sized for: the measured optimum `best` (fewer, fatter streams beat full residency once every warp has
// a batch of loads in flight: INT4 symmetric 0.85 -> 0.92 of the HBM roofline at 3 instead of 8, NVFP4 0.83 -> 0.90 at 4), never more
// than the kernel's own residency `per`.  B200Q_DECODE_CTAS overrides `best` (development sweeps).
int ctas_cap(int per, int best = 0) {
    static const int env = [] { const char* v = getenv("B200Q_DECODE_CTAS"); return (v && *v) ? atoi(v) : 0; }();
    const int want = env > 0 ? env : best;
    return want > 0 ? min(want, per) : per;
}
}  // namespace

// bf16, 4 bits, GROUP (group % 8 == 0, cols % group == 0) or CHANNEL (group == 0).  B200Q_ENOSYS otherwise.
int launch_decode_int4_fast(const int32_t* packed, const void* scale, const int32_t* zp_packed, int64_t batch, int64_t rows, int64_t cols, int group,
                            void* out, cudaStream_t st) {
    if (cols % 8 != 0 || cols >= (1ll << 31) || (group != 0 && (group % 8 != 0 || cols % group != 0))) return B200Q_ENOSYS;
    if ((((uintptr_t)out) & 15) != 0 || (((uintptr_t)packed) & 3) != 0) return B200Q_ENOSYS;
    const int64_t total = batch * rows;
    if (total * cols == 0) return B200Q_OK;
    const uint32_t* pk = (const uint32_t*)packed;
    const uint16_t* sc = (const uint16_t*)scale;
    const uint32_t* zp = (const uint32_t*)zp_packed;
    const int wpg = group / 8, wpr = (int)(cols >> 3);
    // grid = one full wave of the kernel's own residency: a fixed 8 x SMs grid ran the 48-register asymmetric kernel (5 CTAs per SM) as
    // 1.6 waves
#define B200Q_DECODE_INT4(K, ARG, BEST) do { static const int per = [] { int n = 0; return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, K, 256, 0) == cudaSuccess && n > 0) ? n : 4; }(); \
        const unsigned grid = (unsigned)max((int64_t)1, min((int64_t)kNumSMs * ctas_cap(per, BEST), (total + 7) / 8)); \
        K<<<grid, 256, 0, st>>>(pk, sc, zp, total, rows, (int)cols, ARG, (uint4*)out); } while (0)
    static const bool batched = getenv("B200Q_DECODE_INT4_LEGACY") == nullptr;  // A/B switch
    if (batched && wpg > 0 && (wpg & (wpg - 1)) == 0) {
        const int sh = __builtin_ctz((unsigned)wpg);
        const bool u5 = wpr % 160 == 0;  // rows that split into whole batches of 5 words per lane (2560 columns)
        // CTAs per SM from the sweeps in profiles/r1_decompress.json (three shapes): symmetric 3, asymmetric 5 (U = 5) / 4 (U = 4)
        if (zp) { if (u5) B200Q_DECODE_INT4((decode_int4_batch_kernel<true, 5>), sh, 5); else B200Q_DECODE_INT4((decode_int4_batch_kernel<true, 4>), sh, 4); }
        else { if (u5) B200Q_DECODE_INT4((decode_int4_batch_kernel<false, 5>), sh, 3); else B200Q_DECODE_INT4((decode_int4_batch_kernel<false, 4>), sh, 3); }
    } else if (zp) B200Q_DECODE_INT4(decode_int4_kernel<true>, wpg, 0);
    else B200Q_DECODE_INT4(decode_int4_kernel<false>, wpg, 0);
#undef B200Q_DECODE_INT4
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// bf16 scales and output, e4m3 codes, no zero point, no global scale.  strategy: B200Q_TENSOR / CHANNEL / GROUP / BLOCK.
int launch_decode_fp8_fast(const uint8_t* codes, int64_t rows, int64_t cols, int strategy, int group, int bh, int bw, const void* scale, void* out,
                           cudaStream_t st) {
    if (cols % 8 != 0 || cols >= (1ll << 31) || (((uintptr_t)out) & 15) != 0 || (((uintptr_t)codes) & 7) != 0) return B200Q_ENOSYS;
    if (rows * cols == 0) return B200Q_OK;
    const int cpr = (int)(cols >> 3);
    int rows_per_scale = 1, chunks_per_scale = cpr;
    int64_t stride = 1;
    if (strategy == B200Q_TENSOR) { rows_per_scale = 1; stride = 0; }
    else if (strategy == B200Q_CHANNEL) { stride = 1; }
    else if (strategy == B200Q_GROUP) {
        if (group <= 0 || group % 8 != 0 || cols % group != 0) return B200Q_ENOSYS;
        chunks_per_scale = group / 8; stride = cols / group;
    } else if (strategy == B200Q_BLOCK) {
        if (bh <= 0 || bw <= 0 || bw % 8 != 0) return B200Q_ENOSYS;
        rows_per_scale = bh; chunks_per_scale = bw / 8; stride = (cols + bw - 1) / bw;
    } else return B200Q_ENOSYS;
    static const bool batched = getenv("B200Q_DECODE_FP8_LEGACY") == nullptr;  // A/B switch
    const bool row_scale = chunks_per_scale >= cpr, pow2 = (chunks_per_scale & (chunks_per_scale - 1)) == 0;
#define B200Q_DECODE_FP8(K, SH, BEST) do { static const int per = [] { int n = 0; return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, K, 256, 0) == cudaSuccess && n > 0) ? n : 4; }(); \
        const unsigned grid = (unsigned)max((int64_t)1, min((int64_t)kNumSMs * ctas_cap(per, BEST), (rows + 7) / 8)); \
        K<<<grid, 256, 0, st>>>((const uint2*)codes, (const uint16_t*)scale, rows, (int)cols, rows_per_scale, SH, stride, (uint4*)out); } while (0)
    if (batched && row_scale) B200Q_DECODE_FP8((decode_fp8_batch_kernel<true, 4>), 0, 4);  // per-row scale: 0.90 at 3-4 CTAs per SM, 0.86-0.88 at 8
    else if (batched && pow2) B200Q_DECODE_FP8((decode_fp8_batch_kernel<false, 4>), __builtin_ctz((unsigned)chunks_per_scale), 0);
    else B200Q_DECODE_FP8(decode_fp8_kernel, chunks_per_scale, 0);
#undef B200Q_DECODE_FP8
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// bf16 output; gs: fp32 per matrix (gs_stride 1) or shared (0)
int launch_decode_nvfp4_fast(const uint8_t* packed, const uint8_t* scale, const float* gs, int gs_stride, int64_t batch, int64_t rows, int64_t cols,
                             void* out, cudaStream_t st) {
    if (cols % 16 != 0 || (((uintptr_t)out) & 15) != 0 || (((uintptr_t)packed) & 3) != 0 || batch > 65535) return B200Q_ENOSYS;
    const int64_t words = rows * (cols >> 3);
    if (batch * words == 0) return B200Q_OK;
    if ((words * 4) % 4 != 0) return B200Q_ENOSYS;
    // at most ONE wave of resident CTAs over all matrices (the former "+ 1" put `batch` CTAs into a second wave: every CTA owns an equal
    // share of the tiles, so the stragglers doubled the critical path)
    static const int per = [] { int n = 0; return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, decode_nvfp4_kernel, 256, 0) == cudaSuccess && n > 0) ? n : 4; }();
    const unsigned gx = (unsigned)max((int64_t)1, min((int64_t)kNumSMs * ctas_cap(per, 4) / max(batch, (int64_t)1), (words + 1023) / 1024));
    decode_nvfp4_kernel<<<dim3(gx, (unsigned)batch), 256, 0, st>>>((const uint32_t*)packed, scale, gs, gs_stride, words, (uint4*)out);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace b200q
