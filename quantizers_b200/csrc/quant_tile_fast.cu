// quant_tile_fast.cu -- issue-tuned bf16 kernels for the two remaining headline formats:
//   * FP8 128x128 BLOCK (FP8_BLOCK preset): one CTA per tile, the tile lives in registers (8 x 128-bit loads in
//     flight per thread) between the |.|max reduction and the quantization -> one HBM read;
//   * NVFP4 (e2m1 codes + e4m3 per-16 scales): a thread owns whole groups of 16 (two 128-bit loads), so there is
//     no shuffle and the per-group scale chain runs once per group.
// Both use the bracketed reciprocal of fastmath.cuh for the reference's divisions and fall back to the exact IEEE
// chain (qmath.cuh) for the rare elements whose bracket ends disagree.
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.cuh"

namespace b200q {
namespace {
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

template <bool ADD_ZP>
__global__ void __launch_bounds__(256, 3) block_fp8_fast_kernel(const TileParams p) {
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
    br.init(s);
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
            const f32x2 x = bf16x2_to_f32x2(w[k]);
            float al, ah, bl, bh;
            unpack2(ADD_ZP ? mul2_plus0(x, br.lo) : mul2(x, br.lo), al, ah);
            unpack2(ADD_ZP ? mul2_plus0(x, br.hi) : mul2(x, br.hi), bl, bh);
            const uint32_t v = cvt_bf16x2(ah, al);
            diff |= v ^ cvt_bf16x2(bh, bl);
            h[k] = cvt_e4m3x2(__uint_as_float(v & 0xffff0000u), __uint_as_float(v << 16));
        }
        uint2 packed = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
        if (diff != 0 || unsafe) packed = fix_chunk_fp8(raw[i], s, ADD_ZP, unsafe, packed);
        stg_stream((uint8_t*)p.out + mat + r * p.cols + c, packed);
    }
}

// ------------------------------------------------------------------------------------------------ NVFP4
// thread = U2 groups of 16 contiguous elements; warp = 512 contiguous columns per step; CTA = 8 rows
constexpr int U2 = 2;

// fast path: reciprocal normal (s_eff >= 2^-100) and no quotient of a non-zero bf16 (>= 2^-133) underflows to zero
// (s_eff <= 2^16): an underflowed -0.0 would keep its sign through the fused "+ 0.0", the reference's two-step
// (divide, then add the zero-point) turns it into +0.0.
__device__ __forceinline__ bool fp4_scale_is_safe(float s_eff) { return s_eff >= 7.8886090522101181e-31f && s_eff <= 65536.0f; }

__device__ __noinline__ uint32_t fix_group_fp4(const uint4 raw, float s_eff, uint32_t packed) {
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

}  // namespace

int launch_block_fp8_fast(const TileParams& p, int64_t batch, cudaStream_t st) {
    if (p.cols % 8 != 0 || (((uintptr_t)p.w) & 15) != 0 || batch * p.rows * p.cols == 0) return B200Q_ENOSYS;
    dim3 grid((unsigned)((p.cols + 127) / 128), (unsigned)((p.rows + 127) / 128), (unsigned)batch);
    if (grid.y > 65535 || grid.z > 65535) return B200Q_ENOSYS;
    if (p.has_zp) block_fp8_fast_kernel<true><<<grid, 256, 0, st>>>(p);
    else block_fp8_fast_kernel<false><<<grid, 256, 0, st>>>(p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

int launch_nvfp4_fast(const GroupParams& p, int64_t batch, cudaStream_t st) {
    if (p.cols % 16 != 0 || (((uintptr_t)p.w) & 15) != 0 || batch * p.rows * p.cols == 0) return B200Q_ENOSYS;
    if (batch > 65535 || (p.rows + 7) / 8 > 65535) return B200Q_ENOSYS;
    dim3 grid((unsigned)((p.cols + 512 * U2 - 1) / (512 * U2)), (unsigned)((p.rows + 7) / 8), (unsigned)batch);
    nvfp4_fast_kernel<<<grid, 256, 0, st>>>(p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace b200q
