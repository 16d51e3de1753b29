// awq_stats.cu -- AWQ statistics (LLMC modifiers/awq/base.py, restated in SURVEY.md Appendix A):
//   abs_sum_cols   _accumulate_mean numerator     sum_t |x[t,k]|
//   wmean          _compute_layer_means numerator sum_rows |w| / (chunk_absmax + 1e-6), fp64 accumulate
//   awq_scales     _compute_best_scale grid point s = x_mean^r / (w_mean^(1-r) + 1e-4) ... / sqrt(max*min)
//   sq_err         _compute_loss partial          sum (T(y_ref - y_q))^2 in fp32
// All HBM-bound single-pass reductions (2 B/element for bf16).
#include <cstdlib>
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.cuh"

namespace b200q {

// block (32, 8): x -> 256 columns (8 per thread), y -> row lanes; grid (ceil(K/256), row splits)
template <int DT>
__global__ void __launch_bounds__(256) abs_sum_cols_kernel(const void* __restrict__ x, int64_t tokens, int64_t k,
                                                           float* __restrict__ acc) {
    __shared__ float sm[8][256 + 8];
    const int64_t c0 = ((int64_t)blockIdx.x * 32 + threadIdx.x) * 8;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 < k) {
        for (int64_t t = (int64_t)blockIdx.y * 8 + threadIdx.y; t < tokens; t += (int64_t)gridDim.y * 8) {
            Chunk8<DT> ch;
            float v[8];
            load_chunk<DT>(ch, x, t * k + c0);
            chunk_to_float<DT>(ch, v);
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] += fabsf(v[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) sm[threadIdx.y][threadIdx.x * 8 + i] = s[i];
    __syncthreads();
    const int tid = threadIdx.y * 32 + threadIdx.x;  // 256 threads -> 256 columns
    float tot = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; r++) tot += sm[r][tid];
    const int64_t c = (int64_t)blockIdx.x * 256 + tid;
    if (c < k) atomicAdd(&acc[c], tot);
}

int launch_abs_sum_cols(int dt, const void* x, int64_t tokens, int64_t k, float* acc, cudaStream_t st) {
    B200Q_REQUIRE(k % 8 == 0, "hidden size must be a multiple of 8, got %lld", (long long)k);
    B200Q_REQUIRE(((uintptr_t)x & 15) == 0, "activation pointer must be 16-byte aligned");
    if (tokens * k == 0) return B200Q_OK;  // unrouted expert: calibrate_activations skips empty inputs
    const int64_t gx = (k + 255) / 256;
    const int64_t gy = max((int64_t)1, min((tokens + 63) / 64, (int64_t)(kNumSMs * 4 + gx - 1) / gx));
    B200Q_DISPATCH_DT(dt, { abs_sum_cols_kernel<DT><<<dim3((unsigned)gx, (unsigned)gy), dim3(32, 8), 0, st>>>(x, tokens, k, acc); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// one warp per row of the tile, 8 warps stride over rows; lanes own 8 columns of a 256-column tile
template <int DT>
__global__ void __launch_bounds__(256) wmean_kernel(const void* __restrict__ w, int64_t rows, int64_t cols, int group,
                                                    double* __restrict__ acc) {
    __shared__ double sm[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * 256 + lane * 8;
    const int L = group >> 3;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t r_per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r_begin = (int64_t)blockIdx.y * r_per, r_end = min(rows, r_begin + r_per);
    for (int64_t r = r_begin + warp; r < r_begin + ((r_per + 7) / 8) * 8; r += 8) {  // warp-uniform trip count
        const bool ok = r < r_end && c0 < cols;
        float v[8];
        if (ok) {
            Chunk8<DT> ch;
            load_chunk<DT>(ch, w, r * cols + c0);
            chunk_to_float<DT>(ch, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = 0.0f;
        }
        float a = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) { v[i] = fabsf(v[i]); a = fmaxf(a, v[i]); }
        a = subwarp_max(a, L);
        const float den = round_to<DT>(fadd(a, 1e-6f));
        if (ok) {
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] += (double)round_to<DT>(fdiv(v[i], den));
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) sm[warp][lane * 8 + i] = s[i];
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int r = 0; r < 8; r++) tot += sm[r][threadIdx.x];
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < cols) atomicAdd(&acc[c], tot);
}

// bf16 fast path of the same statistic.  The generic kernel above issues one load per warp per row (shuffle inside the loop), an IEEE
// division and an F2F.F64 conversion per element: 0.18 of the HBM roofline.  Here: U rows per warp per batch with every load issued
// first; |max| as packed bf16x2; the quotient T(|w| / den) from the bracketed reciprocal (fastmath.cuh -- both bracket ends round
// to the same bf16 or the chunk is redone with the IEEE division); bf16 -> fp64 by bit placement instead of a conversion.
__device__ __forceinline__ double bf16_nonneg_to_double(uint32_t q) {
    const uint32_t e = q & 0x7f80u;
    if (e == 0u || e == 0x7f80u) return (double)__uint_as_float(q << 16);  // zero, subnormal, inf / NaN
    return __hiloint2double((int)((q << 13) + 0x38000000u), 0);            // rebias 127 -> 1023, mantissa to the top of the 52 bits
}
__device__ __noinline__ void wmean_exact_chunk(const uint4 raw, float den, uint32_t q2[4]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const float lo = fabsf(__uint_as_float(w[k] << 16)), hi = fabsf(__uint_as_float(w[k] & 0xffff0000u));
        q2[k] = fast::cvt_bf16x2(__fdiv_rn(hi, den), __fdiv_rn(lo, den));
    }
}
template <int L, int U>
__global__ void __launch_bounds__(256) wmean_bf16_kernel(const uint16_t* __restrict__ w, int64_t rows, int64_t cols, double* __restrict__ acc) {
    __shared__ double sm[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * 256 + lane * 8;
    const bool col_ok = c0 < cols;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t r_per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r_begin = (int64_t)blockIdx.y * r_per, r_end = min(rows, r_begin + r_per);
    const int64_t r_stop = r_begin + ((r_per + 8 * U - 1) / (8 * U)) * (8 * U);  // warp-uniform trip count (shuffles inside)
    for (int64_t r = r_begin + warp; r < r_stop; r += 8 * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t rr = r + 8 * u;
            v[u] = (rr < r_end && col_ok) ? ldg_stream(w + rr * cols + c0) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = (r + 8 * u) < r_end && col_ok;
            uint32_t a2 = fast::hmaxabs2(fast::hmaxabs2(v[u].x, v[u].y), fast::hmaxabs2(v[u].z, v[u].w));
            a2 = fast::hmaxabs2(a2, fast::prmt(a2, a2, 0x1032));
#pragma unroll
            for (int o = L >> 1; o > 0; o >>= 1) a2 = fast::hmaxabs2(a2, __shfl_xor_sync(0xffffffffu, a2, o));
            const float a = __uint_as_float((a2 << 16) & 0x7fff0000u);
            const float den = round_to<DT_BF16>(fadd(a, 1e-6f));
            fast::Bracket br;
            br.init(den);
            const uint32_t wd[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            uint32_t q2[4], diff = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const fast::f32x2 x = fast::bf16x2_to_f32x2(wd[k] & 0x7fff7fffu);
                float al, ah, bl, bh;
                fast::unpack2(fast::mul2(x, br.lo), al, ah);
                fast::unpack2(fast::mul2(x, br.hi), bl, bh);
                q2[k] = fast::cvt_bf16x2(ah, al);
                diff |= q2[k] ^ fast::cvt_bf16x2(bh, bl);
            }
            if (diff != 0 || !fast::scale_is_safe(__float_as_uint(den))) wmean_exact_chunk(v[u], den, q2);
            if (ok) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    s[2 * k] += bf16_nonneg_to_double(q2[k] & 0xffffu);
                    s[2 * k + 1] += bf16_nonneg_to_double(q2[k] >> 16);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) sm[warp][lane * 8 + i] = s[i];
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int r = 0; r < 8; r++) tot += sm[r][threadIdx.x];
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < cols) atomicAdd(&acc[c], tot);
}

int launch_wmean(int dt, const void* w, int64_t rows, int64_t cols, int group, double* acc, cudaStream_t st) {
    B200Q_REQUIRE(group == 16 || group == 32 || group == 64 || group == 128 || group == 256,
                  "w_mean: group_size %d unsupported (16/32/64/128/256)", group);
    B200Q_REQUIRE(cols % group == 0, "columns %lld not divisible by group_size %d", (long long)cols, group);
    B200Q_REQUIRE(((uintptr_t)w & 15) == 0, "weight pointer must be 16-byte aligned");
    if (rows * cols == 0) return B200Q_OK;
    const int64_t gx = (cols + 255) / 256;
    const int64_t gy = max((int64_t)1, min((rows + 31) / 32, (int64_t)(kNumSMs * 4 + gx - 1) / gx));
    static const bool fast_bf16 = getenv("B200Q_WMEAN_LEGACY") == nullptr;  // A/B switch
    if (fast_bf16 && dt == DT_BF16) {
        const uint16_t* wp = (const uint16_t*)w;
        // row slabs sized so that the whole grid is ONE wave of the kernel's residency (64 registers: 4 CTAs per SM; 600 CTAs on 592
        // slots ran as two waves)
#define B200Q_WMEAN(LANES) do { static const int per = [] { int n = 0; return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wmean_bf16_kernel<LANES, 4>, 256, 0) == cudaSuccess && n > 0) ? n : 2; }(); \
            const int64_t slabs = max((int64_t)1, min((rows + 31) / 32, (int64_t)kNumSMs * per / gx)); \
            wmean_bf16_kernel<LANES, 4><<<dim3((unsigned)gx, (unsigned)slabs), 256, 0, st>>>(wp, rows, cols, acc); } while (0)
        switch (group) {
            case 16: B200Q_WMEAN(2); break;
            case 32: B200Q_WMEAN(4); break;
            case 64: B200Q_WMEAN(8); break;
            case 128: B200Q_WMEAN(16); break;
            default: B200Q_WMEAN(32); break;
        }
#undef B200Q_WMEAN
        B200Q_CHECK_LAUNCH();
        return B200Q_OK;
    }
    B200Q_DISPATCH_DT(dt, { wmean_kernel<DT><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, st>>>(w, rows, cols, group, acc); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// one CTA per grid point; the ratios travel by value in the kernel parameters (no staging allocation, no host copy to order)
struct RatioPack { float r[64]; };
__global__ void __launch_bounds__(256) awq_scales_kernel(const float* __restrict__ x_mean, const float* __restrict__ w_mean,
                                                         int64_t k, const RatioPack ratios, int first, int duo,
                                                         float* __restrict__ scales) {
    __shared__ float smx[8], smn[8];
    const float r = ratios.r[blockIdx.x];
    scales += (int64_t)first * k;
    float* out = scales + (int64_t)blockIdx.x * k;
    float mx = -INFINITY, mn = INFINITY;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        float s = powf(x_mean[i], r);
        if (duo) s = s / (powf(w_mean[i], 1.0f - r) + 1e-4f);
        s = fmaxf(s, 1e-4f);  // clamp(min=1e-4); NaN stays NaN in torch, fmaxf drops it -> fixed below by the isnan rule
        out[i] = s;
        mx = fmaxf(mx, s);
        mn = fminf(mn, s);
    }
    mx = subwarp_max(mx, 32);
    mn = subwarp_min(mn, 32);
    if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; }
    __syncthreads();
    mx = smx[0]; mn = smn[0];
    for (int i = 1; i < 8; i++) { mx = fmaxf(mx, smx[i]); mn = fminf(mn, smn[i]); }
    const float norm = sqrtf(mx * mn);
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        float s = out[i] / norm;
        if (isinf(s) || isnan(s)) s = 1.0f;
        out[i] = s;
    }
}

int launch_awq_scales(const float* x_mean, const float* w_mean, int64_t k, const float* ratios, int n_ratios, int duo,
                      float* scales, cudaStream_t st) {
    if (k == 0 || n_ratios == 0) return B200Q_OK;
    B200Q_REQUIRE(!duo || w_mean != nullptr, "duo_scaling needs w_mean");
    for (int first = 0; first < n_ratios; first += 64) {  // ratios: HOST pointer
        RatioPack pack;
        const int n = n_ratios - first < 64 ? n_ratios - first : 64;
        for (int i = 0; i < n; i++) pack.r[i] = ratios[first + i];
        awq_scales_kernel<<<n, 256, 0, st>>>(x_mean, w_mean, k, pack, first, duo, scales);
        B200Q_CHECK_LAUNCH();
    }
    return B200Q_OK;
}

// ---- routed MoE block output (W2 of the layer-wide MoE mapping): out[t] = sum_j bf16(y[row[t, j]] * w[t, j]) accumulated in bf16,
// j in ascending expert order -- exactly the rounding sequence of transformers' per-expert ``index_add_`` into a bf16 tensor.
// `init` (may alias `out`): the running sum of the experts that come BEFORE this call's in that order (expert-parallel ring: rank r
// continues the sequence rank r - 1 left off); null = the reference's zero-initialised output.
__global__ void __launch_bounds__(256) moe_combine_kernel(const uint16_t* __restrict__ y, const int32_t* __restrict__ row, const uint16_t* __restrict__ w,
                                                          int64_t tokens, int top_k, int64_t h, const uint16_t* init, uint16_t* out) {
    const int64_t chunks = h / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tokens * chunks) return;
    const int64_t t = i / chunks, c = i - t * chunks;
    float acc[8];
    if (init != nullptr) {
        Chunk8<DT_BF16> ch;
        load_chunk<DT_BF16>(ch, init, t * h + c * 8);
        chunk_to_float<DT_BF16>(ch, acc);
    } else {
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = 0.0f;
    }
    for (int j = 0; j < top_k; j++) {
        const int32_t r = row[t * top_k + j];
        if (r < 0) continue;
        const float wf = __uint_as_float((uint32_t)w[t * top_k + j] << 16);
        Chunk8<DT_BF16> ch;
        float v[8];
        load_chunk<DT_BF16>(ch, y, (int64_t)r * h + c * 8);
        chunk_to_float<DT_BF16>(ch, v);
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = round_to<DT_BF16>(fadd(acc[e], round_to<DT_BF16>(fmul(v[e], wf))));
    }
    store_chunk_T<DT_BF16>(out, t * h + c * 8, acc);
}
int launch_moe_combine(const void* y, const int32_t* row, const void* w, int64_t tokens, int top_k, int64_t h, const void* init, void* out,
                       cudaStream_t st) {
    B200Q_REQUIRE(h % 8 == 0 && top_k >= 1, "moe_combine: hidden size must be a multiple of 8");
    if (tokens == 0) return B200Q_OK;
    const int64_t n = tokens * (h / 8);
    moe_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint16_t*)y, row, (const uint16_t*)w, tokens, top_k, h,
                                                                    (const uint16_t*)init, (uint16_t*)out);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

template <int DT>
__global__ void __launch_bounds__(256) sq_err_kernel(const void* __restrict__ a, const void* __restrict__ b, int64_t n,
                                                     float* __restrict__ acc) {
    __shared__ float sm[8];
    float s = 0.0f;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n / 8; t += (int64_t)gridDim.x * blockDim.x) {
        Chunk8<DT> ca, cb;
        float x[8], y[8];
        load_chunk<DT>(ca, a, t * 8);
        load_chunk<DT>(cb, b, t * 8);
        chunk_to_float<DT>(ca, x);
        chunk_to_float<DT>(cb, y);
#pragma unroll
        for (int i = 0; i < 8; i++) { const float d = round_to<DT>(fadd(x[i], -y[i])); s = fmaf(d, d, s); }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n % 8)) {
        const float d = round_to<DT>(fadd(load_T<DT>(a, (n / 8) * 8 + threadIdx.x), -load_T<DT>(b, (n / 8) * 8 + threadIdx.x)));
        s = fmaf(d, d, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int i = 0; i < 8; i++) tot += sm[i];
        atomicAdd(acc, tot);
    }
}

int launch_sq_err(int dt, const void* a, const void* b, int64_t n, float* acc, cudaStream_t st) {
    B200Q_REQUIRE((((uintptr_t)a | (uintptr_t)b) & 15) == 0, "output pointers must be 16-byte aligned");
    if (n == 0) return B200Q_OK;
    const unsigned g = (unsigned)max((int64_t)1, min((n / 8 + 255) / 256, (int64_t)kNumSMs * 8));
    B200Q_DISPATCH_DT(dt, { sq_err_kernel<DT><<<g, 256, 0, st>>>(a, b, n, acc); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace b200q
