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
