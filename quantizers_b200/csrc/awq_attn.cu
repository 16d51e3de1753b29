// awq_attn.cu -- the element-wise part of the attention parent in AWQ's _run_samples (LLMC modifiers/awq/base.py; parent =
// transformers Qwen3Attention for the input_layernorm -> q/k/v mapping): per-head RMSNorm of q and k followed by the rotary
// embedding, in place on the projected [tokens, (H + 2 Hkv) d] rows that awq_gemm_project_kernel wrote.  The reference does this
// with ~14 eager ATen passes (pow, mean, rsqrt, mul, cat, neg, ...) per forward; here it is one HBM read + one write of the q/k
// columns (v is untouched).  Every rounding point of the bf16 eager chain is reproduced:
//     vn  = bf16(x * rsqrt(mean(x^2) + eps))      (fp32 inside, rounded once)
//     y   = bf16(w * vn)
//     out = bf16( bf16(y * cos) + bf16(rotate_half(y) * sin) )
// One warp per (token, head): a lane owns d/32 consecutive elements; rotate_half's partner (i +- d/2) lives in lane ^ 16.
#include <cstdlib>
#include "../../include/b200q.h"
#include "common.cuh"

namespace b200q {
namespace {

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }

// one vector per warp per step (default)
template <int EPL>  // elements per lane: 2 (d = 64) or 4 (d = 128)
__global__ void __launch_bounds__(256) qk_norm_rope_v1_kernel(uint16_t* __restrict__ qkv, int64_t tokens, int n_heads, int n_kv, int seq_len,
                                                           const uint16_t* __restrict__ qw, const uint16_t* __restrict__ kw,
                                                           const uint16_t* __restrict__ cosb, const uint16_t* __restrict__ sinb, float eps) {
    constexpr int D = EPL * 32;
    const int lane = threadIdx.x & 31;
    const int64_t n_vec = tokens * (int64_t)(n_heads + n_kv);
    const int64_t row_elems = (int64_t)(n_heads + 2 * n_kv) * D;
    for (int64_t vec = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); vec < n_vec; vec += (int64_t)gridDim.x * 8) {
        const int64_t t = vec / (n_heads + n_kv);
        const int h = (int)(vec - t * (n_heads + n_kv));  // heads 0..H-1 are q, H..H+Hkv-1 are k
        const uint16_t* nw = h < n_heads ? qw : kw;
        uint16_t* p = qkv + t * row_elems + (int64_t)h * D + lane * EPL;
        const int pos = (int)(t % seq_len);
        uint32_t raw[EPL / 2], wr[EPL / 2], cr[EPL / 2], sr[EPL / 2];
        if (EPL == 4) {
            const uint2 a = *reinterpret_cast<const uint2*>(p);
            raw[0] = a.x; raw[EPL / 2 - 1] = a.y;
            const uint2 b = *reinterpret_cast<const uint2*>(nw + lane * EPL);
            wr[0] = b.x; wr[EPL / 2 - 1] = b.y;
            const uint2 c = *reinterpret_cast<const uint2*>(cosb + (int64_t)pos * D + lane * EPL);
            cr[0] = c.x; cr[EPL / 2 - 1] = c.y;
            const uint2 s = *reinterpret_cast<const uint2*>(sinb + (int64_t)pos * D + lane * EPL);
            sr[0] = s.x; sr[EPL / 2 - 1] = s.y;
        } else {
            raw[0] = *reinterpret_cast<const uint32_t*>(p);
            wr[0] = *reinterpret_cast<const uint32_t*>(nw + lane * EPL);
            cr[0] = *reinterpret_cast<const uint32_t*>(cosb + (int64_t)pos * D + lane * EPL);
            sr[0] = *reinterpret_cast<const uint32_t*>(sinb + (int64_t)pos * D + lane * EPL);
        }
        float x[EPL], ss = 0.0f;
#pragma unroll
        for (int i = 0; i < EPL; i++) {
            x[i] = bf16_bits_to_float((raw[i >> 1] >> (16 * (i & 1))) & 0xffffu);
            ss = __fadd_rn(ss, __fmul_rn(x[i], x[i]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, o));
        const float r = rsqrtf(__fadd_rn(__fdiv_rn(ss, (float)D), eps));
        float y[EPL], out[EPL];
#pragma unroll
        for (int i = 0; i < EPL; i++) {
            const float w = bf16_bits_to_float((wr[i >> 1] >> (16 * (i & 1))) & 0xffffu);
            y[i] = bf16r(__fmul_rn(w, bf16r(__fmul_rn(x[i], r))));
        }
#pragma unroll
        for (int i = 0; i < EPL; i++) {
            const float partner = __shfl_xor_sync(0xffffffffu, y[i], 16);
            const float rot = lane < 16 ? -partner : partner;  // first half: -x[i + d/2]; second half: x[i - d/2]
            const float c = bf16_bits_to_float((cr[i >> 1] >> (16 * (i & 1))) & 0xffffu);
            const float s = bf16_bits_to_float((sr[i >> 1] >> (16 * (i & 1))) & 0xffffu);
            out[i] = __fadd_rn(bf16r(__fmul_rn(y[i], c)), bf16r(__fmul_rn(rot, s)));
        }
        uint32_t o2[EPL / 2];
#pragma unroll
        for (int i = 0; i < EPL / 2; i++) {
            __nv_bfloat162 hv = __floats2bfloat162_rn(out[2 * i], out[2 * i + 1]);
            o2[i] = *reinterpret_cast<uint32_t*>(&hv);
        }
        if (EPL == 4) *reinterpret_cast<uint2*>(p) = make_uint2(o2[0], o2[EPL / 2 - 1]);
        else *reinterpret_cast<uint32_t*>(p) = o2[0];
    }
}

// v2 (default, round 2): D / 16 lanes per (token, head) vector.  A lane owns elements [8l, 8l + 8) AND [D/2 + 8l, D/2 + 8l + 8) -- two 16-byte
// loads, still 128 contiguous bytes per half across the lanes -- so rotate_half's partner (i +- D/2) is in the SAME thread: no
// shuffles for the rotation, log2(D / 16) for mean(x^2).  The bf16 chain runs packed: vn = cvt.rn.bf16x2(x * r), y = w * vn,
// y * cos, rot * sin (HMUL2: the exact product rounded once, what the eager bf16 multiply gives), their sum one HADD2.
// ~6 issued instructions per element instead of ~30 (v1: 4 elements per lane, 9 shuffles per 4 elements, scalar chain).
__device__ __forceinline__ uint32_t rope_hmul2(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t rope_hadd2(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t rope_cvt2(float hi, float lo) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int D>
__global__ void __launch_bounds__(256) qk_norm_rope_v2_kernel(uint16_t* __restrict__ qkv, int64_t tokens, int n_heads, int n_kv, int seq_len,
                                                           const uint16_t* __restrict__ qw, const uint16_t* __restrict__ kw,
                                                           const uint16_t* __restrict__ cosb, const uint16_t* __restrict__ sinb, float eps) {
    constexpr int LPV = D / 16;        // lanes per vector: 8 (d = 128) or 4 (d = 64)
    constexpr int VPW = 32 / LPV;      // vectors per warp
    const int lane = threadIdx.x & 31, sub = lane % LPV, vslot = lane / LPV;
    const int hq = n_heads + n_kv;
    const int64_t n_vec = tokens * (int64_t)hq;
    const int64_t row_elems = (int64_t)(n_heads + 2 * n_kv) * D;
    // norm weights of this lane's 16 elements, q and k (loaded once)
    const uint4 wq0 = *reinterpret_cast<const uint4*>(qw + sub * 8), wq1 = *reinterpret_cast<const uint4*>(qw + D / 2 + sub * 8);
    const uint4 wk0 = *reinterpret_cast<const uint4*>(kw + sub * 8), wk1 = *reinterpret_cast<const uint4*>(kw + D / 2 + sub * 8);
    const int64_t wstride = (int64_t)gridDim.x * 8 * VPW;
    for (int64_t vbase = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * VPW; vbase < n_vec; vbase += wstride) {   // warp-uniform
        const int64_t vec = vbase + vslot;
        const bool ok = vec < n_vec;
        const int64_t vv = ok ? vec : n_vec - 1;   // idle sub-warps recompute the last vector and store nothing (shuffles stay full-warp)
        const int64_t t = vv / hq;
        const int h = (int)(vv - t * hq);
        const bool isq = h < n_heads;
        uint16_t* p = qkv + t * row_elems + (int64_t)h * D + sub * 8;
        const int64_t pos = (t % seq_len) * D + sub * 8;
        const uint4 a0 = *reinterpret_cast<const uint4*>(p), a1 = *reinterpret_cast<const uint4*>(p + D / 2);
        const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(cosb + pos)), c1 = __ldg(reinterpret_cast<const uint4*>(cosb + pos + D / 2));
        const uint4 s0 = __ldg(reinterpret_cast<const uint4*>(sinb + pos)), s1 = __ldg(reinterpret_cast<const uint4*>(sinb + pos + D / 2));
        const uint32_t x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const uint4 w0 = isq ? wq0 : wk0, w1 = isq ? wq1 : wk1;
        const uint32_t nw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        float xl[8], xh[8], ss0 = 0.0f, ss1 = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            xl[i] = __uint_as_float(x[i] << 16);
            xh[i] = __uint_as_float(x[i] & 0xffff0000u);
            ss0 = __fadd_rn(ss0, __fmul_rn(xl[i], xl[i]));
            ss1 = __fadd_rn(ss1, __fmul_rn(xh[i], xh[i]));
        }
        float ss = __fadd_rn(ss0, ss1);
#pragma unroll
        for (int o = 1; o < LPV; o <<= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, o));
        const float r = rsqrtf(__fadd_rn(__fdiv_rn(ss, (float)D), eps));
        uint32_t y[8];
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = rope_hmul2(nw[i], rope_cvt2(__fmul_rn(xh[i], r), __fmul_rn(xl[i], r)));   // w * T(x * r)
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // first half: y * cos + (-y[i + D/2]) * sin ; second half: y * cos + y[i - D/2] * sin
            o[i] = rope_hadd2(rope_hmul2(y[i], cw[i]), rope_hmul2(y[i + 4] ^ 0x80008000u, sw[i]));
            o[i + 4] = rope_hadd2(rope_hmul2(y[i + 4], cw[i + 4]), rope_hmul2(y[i], sw[i + 4]));
        }
        if (ok) {
            *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(p + D / 2) = make_uint4(o[4], o[5], o[6], o[7]);
        }
    }
}

// Experiment (B200Q_ROPE_BATCH=4): U (token, head) vectors per warp per step, the loads of all U vectors (row, cos, sin) issued before the
// first shuffle; the arithmetic per vector is unchanged.  Measured SLOWER than the one-vector schedule (365 vs 323 us): not latency-bound.
template <int EPL, int U>  // EPL: elements per lane: 2 (d = 64) or 4 (d = 128)
__global__ void __launch_bounds__(256) qk_norm_rope_kernel(uint16_t* __restrict__ qkv, int64_t tokens, int n_heads, int n_kv, int seq_len,
                                                           const uint16_t* __restrict__ qw, const uint16_t* __restrict__ kw,
                                                           const uint16_t* __restrict__ cosb, const uint16_t* __restrict__ sinb, float eps) {
    constexpr int D = EPL * 32;
    constexpr int W = EPL / 2;  // 32-bit words per lane
    const int lane = threadIdx.x & 31;
    const int hq = n_heads + n_kv;
    const int64_t n_vec = tokens * (int64_t)hq;
    const int64_t row_elems = (int64_t)(n_heads + 2 * n_kv) * D;
    const int64_t vstride = (int64_t)gridDim.x * 8;
    // the norm weights depend only on (q or k, lane): both loaded once
    uint32_t wq[W], wk[W];
#pragma unroll
    for (int i = 0; i < W; i++) {
        wq[i] = *reinterpret_cast<const uint32_t*>(qw + lane * EPL + 2 * i);
        wk[i] = *reinterpret_cast<const uint32_t*>(kw + lane * EPL + 2 * i);
    }
    for (int64_t vec0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); vec0 < n_vec; vec0 += vstride * U) {  // warp-uniform
        uint32_t raw[U][W], cr[U][W], sr[U][W];
        uint16_t* ptr[U];
        bool isq[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t vec = min(vec0 + u * vstride, n_vec - 1);  // out-of-range slots recompute the last vector and store nothing
            const int64_t t = vec / hq;
            const int h = (int)(vec - t * hq);  // heads 0..H-1 are q, H..H+Hkv-1 are k
            isq[u] = h < n_heads;
            ptr[u] = qkv + t * row_elems + (int64_t)h * D + lane * EPL;
            const int64_t pos = (t % seq_len) * D + lane * EPL;
#pragma unroll
            for (int i = 0; i < W; i++) {
                raw[u][i] = *reinterpret_cast<const uint32_t*>(ptr[u] + 2 * i);
                cr[u][i] = __ldg(reinterpret_cast<const uint32_t*>(cosb + pos + 2 * i));
                sr[u][i] = __ldg(reinterpret_cast<const uint32_t*>(sinb + pos + 2 * i));
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            float x[EPL], ss = 0.0f;
#pragma unroll
            for (int i = 0; i < EPL; i++) {
                x[i] = bf16_bits_to_float((raw[u][i >> 1] >> (16 * (i & 1))) & 0xffffu);
                ss = __fadd_rn(ss, __fmul_rn(x[i], x[i]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, o));
            const float r = rsqrtf(__fadd_rn(__fdiv_rn(ss, (float)D), eps));
            float y[EPL], out[EPL];
#pragma unroll
            for (int i = 0; i < EPL; i++) {
                const uint32_t wr = isq[u] ? wq[i >> 1] : wk[i >> 1];
                const float w = bf16_bits_to_float((wr >> (16 * (i & 1))) & 0xffffu);
                y[i] = bf16r(__fmul_rn(w, bf16r(__fmul_rn(x[i], r))));
            }
#pragma unroll
            for (int i = 0; i < EPL; i++) {
                const float partner = __shfl_xor_sync(0xffffffffu, y[i], 16);
                const float rot = lane < 16 ? -partner : partner;  // first half: -x[i + d/2]; second half: x[i - d/2]
                const float c = bf16_bits_to_float((cr[u][i >> 1] >> (16 * (i & 1))) & 0xffffu);
                const float s = bf16_bits_to_float((sr[u][i >> 1] >> (16 * (i & 1))) & 0xffffu);
                out[i] = __fadd_rn(bf16r(__fmul_rn(y[i], c)), bf16r(__fmul_rn(rot, s)));
            }
            if (vec0 + u * vstride < n_vec) {
#pragma unroll
                for (int i = 0; i < W; i++) {
                    __nv_bfloat162 hv = __floats2bfloat162_rn(out[2 * i], out[2 * i + 1]);
                    *reinterpret_cast<uint32_t*>(ptr[u] + 2 * i) = *reinterpret_cast<uint32_t*>(&hv);
                }
            }
        }
    }
}

}  // namespace
}  // namespace b200q

using namespace b200q;

extern "C" int b200q_qk_norm_rope(void* qkv, int64_t tokens, int32_t n_heads, int32_t n_kv, int32_t head_dim, int32_t seq_len,
                                  const void* q_norm_weight, const void* k_norm_weight, const void* cos, const void* sin, float eps,
                                  void* stream) {
    B200Q_REQUIRE(qkv && q_norm_weight && k_norm_weight && cos && sin, "b200q_qk_norm_rope: NULL pointer");
    B200Q_REQUIRE(head_dim == 64 || head_dim == 128, "head_dim must be 64 or 128, got %d", (int)head_dim);
    B200Q_REQUIRE(n_heads >= 1 && n_kv >= 1 && seq_len >= 1 && tokens >= 0, "bad attention geometry");
    B200Q_REQUIRE((((uintptr_t)qkv | (uintptr_t)q_norm_weight | (uintptr_t)k_norm_weight | (uintptr_t)cos | (uintptr_t)sin) & 7) == 0,
                  "pointers must be 8-byte aligned");
    if (tokens == 0) return B200Q_OK;
    const int64_t n_vec = tokens * (int64_t)(n_heads + n_kv);
    cudaStream_t st = (cudaStream_t)stream;
    // default: v2 (sub-warp per vector, packed bf16 chain); B200Q_ROPE_V1=1 / B200Q_ROPE_BATCH=4 select the round-1 kernels (same
    // rounding chain; the mean(x^2) summation order differs, as it does from the eager reference)
    static const bool v1 = getenv("B200Q_ROPE_V1") != nullptr || getenv("B200Q_ROPE_BATCH") != nullptr;
    if (!v1 && (((uintptr_t)qkv | (uintptr_t)q_norm_weight | (uintptr_t)k_norm_weight | (uintptr_t)cos | (uintptr_t)sin) & 15) == 0) {
        const int vpw = head_dim == 128 ? 4 : 8;
        const int grid = (int)min((n_vec + 8 * vpw - 1) / (8 * vpw), (int64_t)kNumSMs * 8);
        if (head_dim == 128)
            qk_norm_rope_v2_kernel<128><<<grid, 256, 0, st>>>((uint16_t*)qkv, tokens, n_heads, n_kv, seq_len, (const uint16_t*)q_norm_weight,
                                                              (const uint16_t*)k_norm_weight, (const uint16_t*)cos, (const uint16_t*)sin, eps);
        else
            qk_norm_rope_v2_kernel<64><<<grid, 256, 0, st>>>((uint16_t*)qkv, tokens, n_heads, n_kv, seq_len, (const uint16_t*)q_norm_weight,
                                                             (const uint16_t*)k_norm_weight, (const uint16_t*)cos, (const uint16_t*)sin, eps);
        B200Q_CHECK_LAUNCH();
        return B200Q_OK;
    }
    // B200Q_ROPE_BATCH=4 selects the four-vectors-per-step schedule (A/B switch; its grid is one wave of the kernel's residency)
    // measured (T = 32 768, 32 + 8 heads of 128): one vector per step 323 us, four per step 365 us -- the kernel is bound by its ~30
    // instructions per element (0.3 of the HBM roofline either way), not by load latency, and 73 registers leave 3 CTAs per SM
    static const int batch = [] { const char* v = getenv("B200Q_ROPE_BATCH"); return (v && v[0] == '4') ? 4 : 1; }();
#define B200Q_ROPE(EPL_, U_) do { \
        static const int per = [] { int n = 0; return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, qk_norm_rope_kernel<EPL_, U_>, 256, 0) == cudaSuccess && n > 0) ? n : 2; }(); \
        const int grid = (int)min((n_vec + 8 * U_ - 1) / (8 * U_), (int64_t)kNumSMs * per); \
        qk_norm_rope_kernel<EPL_, U_><<<grid, 256, 0, st>>>((uint16_t*)qkv, tokens, n_heads, n_kv, seq_len, (const uint16_t*)q_norm_weight, \
                                                        (const uint16_t*)k_norm_weight, (const uint16_t*)cos, (const uint16_t*)sin, eps); } while (0)
#define B200Q_ROPE_V1(EPL_) do { \
        const int grid = (int)min((n_vec + 7) / 8, (int64_t)kNumSMs * 16); \
        qk_norm_rope_v1_kernel<EPL_><<<grid, 256, 0, st>>>((uint16_t*)qkv, tokens, n_heads, n_kv, seq_len, (const uint16_t*)q_norm_weight, \
                                                       (const uint16_t*)k_norm_weight, (const uint16_t*)cos, (const uint16_t*)sin, eps); } while (0)
    if (head_dim == 128) { if (batch == 1) B200Q_ROPE_V1(4); else B200Q_ROPE(4, 4); }
    else { if (batch == 1) B200Q_ROPE_V1(2); else B200Q_ROPE(2, 4); }
#undef B200Q_ROPE_V1
#undef B200Q_ROPE
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}
