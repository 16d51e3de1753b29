// observers.cu -- calibration observers as stand-alone kernels.
//   O1 min/max per quantization chunk  (LLMC observers/helpers.py flatten_for_calibration + min_max.py _get_min_max)
//   O2/O3 per-tensor running min/max + generate_gparam  (LLMC Observer.get_global_scale, static_minmax;
//         CT:quantization/utils/helpers.py:309-338)
//   Q1 calculate_qparams  (CT:quantization/utils/helpers.py:50-137)
// All are single-pass, HBM-bound reads with 128-bit loads; results are exact (min/max are order independent).
#include <cstdlib>
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.cuh"

namespace b200q {

__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
    if (v >= 0.0f) atomicMax((int*)addr, __float_as_int(v));
    else atomicMin((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
    if (v >= 0.0f) atomicMin((int*)addr, __float_as_int(v));
    else atomicMax((unsigned int*)addr, __float_as_uint(v));
}

// ---- GROUP (g = 16..256, sub-warp reduce) and CHANNEL (whole row): one warp per row, 256 columns per step
template <int DT>
__global__ void __launch_bounds__(256) minmax_rows_kernel(const void* __restrict__ w, int64_t nrows, int64_t cols, int group,
                                                          void* __restrict__ mn_out, void* __restrict__ mx_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const bool channel = group <= 0;
    const int L = channel ? 32 : (group >> 3);
    const int64_t gtot = channel ? 1 : cols / group;
    float rmn = INFINITY, rmx = -INFINITY;
    const int64_t nsteps = (cols + 255) / 256;  // warp-uniform trip count (shuffles inside)
    for (int64_t step = 0; step < nsteps; step++) {
        const int64_t c0 = step * 256 + (int64_t)lane * 8;
        const bool ok = c0 < cols;
        float mn = INFINITY, mx = -INFINITY;
        if (ok) {
            Chunk8<DT> ch;
            float x[8];
            load_chunk<DT>(ch, w, row * cols + c0);
            chunk_to_float<DT>(ch, x);
#pragma unroll
            for (int i = 0; i < 8; i++) { mn = fminf(mn, x[i]); mx = fmaxf(mx, x[i]); }
        }
        if (channel) { rmn = fminf(rmn, mn); rmx = fmaxf(rmx, mx); continue; }
        mn = subwarp_min(mn, L);
        mx = subwarp_max(mx, L);
        if (ok && (lane % L) == 0) {
            store_T<DT>(mn_out, row * gtot + c0 / group, mn);
            store_T<DT>(mx_out, row * gtot + c0 / group, mx);
        }
    }
    if (channel) {
        rmn = subwarp_min(rmn, 32);
        rmx = subwarp_max(rmx, 32);
        if (lane == 0) { store_T<DT>(mn_out, row, rmn); store_T<DT>(mx_out, row, rmx); }
    }
}

// ---- bf16 GROUP fast path.  The generic kernel above keeps its shuffles inside the column loop, so ptxas issues ONE 16-byte load per
// warp per step and the kernel is latency-bound (measured 0.50-0.60 of the HBM roofline against 0.88 for CHANNEL, whose loop has no
// shuffle).  Here a warp takes a row in batches of U chunks per lane -- all loads of a batch before the first reduction -- min and max
// are packed bf16x2 instructions (exact: no arithmetic, same NaN-dropping / -0 < +0 rules as FMNMX), and the sub-warp reduction
// carries {min, max} in ONE register (low / high half) so a step costs one SHFL.
template <int L /* lanes per group = group / 8 */, int U>
__global__ void __launch_bounds__(256) minmax_group_bf16_kernel(const uint4* __restrict__ w, int64_t nrows, int cpr /* 16-byte chunks per row */,
                                                                uint16_t* __restrict__ mn_out, uint16_t* __restrict__ mx_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const uint4* rowp = w + row * cpr;
    const int64_t obase = row * (cpr / L);
    for (int base = 0; base < cpr; base += 32 * U) {  // warp-uniform trip count (shuffles inside)
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int c = base + 32 * u + lane;
            v[u] = c < cpr ? ldg_stream(rowp + c) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int c = base + 32 * u + lane;
            uint32_t mn = fast::hmin2(fast::hmin2(v[u].x, v[u].y), fast::hmin2(v[u].z, v[u].w));
            uint32_t mx = fast::hmax2(fast::hmax2(v[u].x, v[u].y), fast::hmax2(v[u].z, v[u].w));
            mn = fast::hmin2(mn, fast::prmt(mn, mn, 0x1032));
            mx = fast::hmax2(mx, fast::prmt(mx, mx, 0x1032));
            uint32_t mm = fast::prmt(mn, mx, 0x5410);  // low half: min, high half: max
#pragma unroll
            for (int o = L >> 1; o > 0; o >>= 1) {
                const uint32_t t = __shfl_xor_sync(0xffffffffu, mm, o);
                mm = (fast::hmin2(mm, t) & 0x0000ffffu) | (fast::hmax2(mm, t) & 0xffff0000u);
            }
            // cols % group == 0: the L lanes of a group are all inside the row or all outside it
            if (c < cpr && (lane & (L - 1)) == 0) {
                const int64_t g = obase + c / L;
                mn_out[g] = (uint16_t)(mm & 0xffffu);
                mx_out[g] = (uint16_t)(mm >> 16);
            }
        }
    }
}

template <int L>
void launch_minmax_group_bf16(const void* w, int64_t nrows, int64_t cols, void* mn, void* mx, cudaStream_t st) {
    minmax_group_bf16_kernel<L, 4><<<(unsigned)((nrows + 7) / 8), 256, 0, st>>>((const uint4*)w, nrows, (int)(cols >> 3), (uint16_t*)mn, (uint16_t*)mx);
}

// ---- BLOCK: one CTA per (block-row, block-col) tile; ragged edges see the zero padding CT applies
template <int DT>
__global__ void __launch_bounds__(256) minmax_block_kernel(const void* __restrict__ w, int64_t rows, int64_t cols, int bh, int bw,
                                                           void* __restrict__ mn_out, void* __restrict__ mx_out) {
    const int64_t b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * bh, c0 = (int64_t)blockIdx.x * bw;
    const char* base = (const char*)w + b * rows * cols * ElemSize<DT>::value;
    float mn = INFINITY, mx = -INFINITY;
    const bool ragged = (r0 + bh > rows) || (c0 + bw > cols);
    if (ragged) { mn = 0.0f; mx = 0.0f; }
    const int cpr = bw / 8;  // chunks per tile row (bw % 8 == 0 enforced by the launcher)
    for (int t = threadIdx.x; t < bh * cpr; t += blockDim.x) {
        const int64_t r = r0 + t / cpr, c = c0 + (int64_t)(t % cpr) * 8;
        if (r < rows && c < cols) {
            Chunk8<DT> ch;
            float x[8];
            load_chunk<DT>(ch, base, r * cols + c);
            chunk_to_float<DT>(ch, x);
#pragma unroll
            for (int i = 0; i < 8; i++) { mn = fminf(mn, x[i]); mx = fmaxf(mx, x[i]); }
        }
    }
    __shared__ float smn[8], smx[8];
    mn = subwarp_min(mn, 32);
    mx = subwarp_max(mx, 32);
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
        const int64_t k = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        store_T<DT>(mn_out, k, mn);
        store_T<DT>(mx_out, k, mx);
    }
}

// bf16 128x128 fast path: the tile is 2048 16-byte chunks = 8 per thread, all loaded before the first reduction (the generic loop
// above divides by a run-time tile width and loads under a branch: one load in flight per thread, 0.76 of the HBM roofline).  Zero
// fill is exact: a chunk is missing only in a ragged tile, where CT's zero padding takes part in the min / max anyway.
__global__ void __launch_bounds__(256) minmax_block128_bf16_kernel(const uint16_t* __restrict__ w, int64_t rows, int64_t cols,
                                                                   uint16_t* __restrict__ mn_out, uint16_t* __restrict__ mx_out) {
    const int64_t b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * 128 + (threadIdx.x >> 4), c = (int64_t)blockIdx.x * 128 + (threadIdx.x & 15) * 8;
    const uint16_t* base = w + b * rows * cols;
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int64_t r = r0 + 16 * j;
        v[j] = (r < rows && c < cols) ? ldg_stream(base + r * cols + c) : make_uint4(0u, 0u, 0u, 0u);
    }
    uint32_t mn = fast::hmin2(fast::hmin2(v[0].x, v[0].y), fast::hmin2(v[0].z, v[0].w));
    uint32_t mx = fast::hmax2(fast::hmax2(v[0].x, v[0].y), fast::hmax2(v[0].z, v[0].w));
#pragma unroll
    for (int j = 1; j < 8; j++) {
        mn = fast::hmin2(mn, fast::hmin2(fast::hmin2(v[j].x, v[j].y), fast::hmin2(v[j].z, v[j].w)));
        mx = fast::hmax2(mx, fast::hmax2(fast::hmax2(v[j].x, v[j].y), fast::hmax2(v[j].z, v[j].w)));
    }
    mn = fast::hmin2(mn, fast::prmt(mn, mn, 0x1032));
    mx = fast::hmax2(mx, fast::prmt(mx, mx, 0x1032));
    uint32_t mm = fast::prmt(mn, mx, 0x5410);  // low half: min, high half: max
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t t = __shfl_xor_sync(0xffffffffu, mm, o);
        mm = (fast::hmin2(mm, t) & 0x0000ffffu) | (fast::hmax2(mm, t) & 0xffff0000u);
    }
    __shared__ uint32_t sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mm;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < 8; i++) mm = (fast::hmin2(mm, sm[i]) & 0x0000ffffu) | (fast::hmax2(mm, sm[i]) & 0xffff0000u);
        const int64_t k = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        mn_out[k] = (uint16_t)(mm & 0xffffu);
        mx_out[k] = (uint16_t)(mm >> 16);
    }
}

// ---- TENSOR: grid-stride reduce into fp32 {min,max} state per batch entry
__global__ void init_minmax_state_kernel(float* state, int64_t batch) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) { state[2 * i] = INFINITY; state[2 * i + 1] = -INFINITY; }
}
template <int DT>
__global__ void __launch_bounds__(256) minmax_tensor_kernel(const void* __restrict__ x, int64_t numel, float* __restrict__ state) {
    const int64_t b = blockIdx.y;
    const char* base = (const char*)x + b * numel * ElemSize<DT>::value;
    float mn = INFINITY, mx = -INFINITY;
    const int64_t nchunks = numel / 8;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nchunks; t += (int64_t)gridDim.x * blockDim.x) {
        Chunk8<DT> ch;
        float v[8];
        load_chunk<DT>(ch, base, t * 8);
        chunk_to_float<DT>(ch, v);
#pragma unroll
        for (int i = 0; i < 8; i++) { mn = fminf(mn, v[i]); mx = fmaxf(mx, v[i]); }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(numel % 8)) {
        const float v = load_T<DT>(base, nchunks * 8 + threadIdx.x);
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    mn = subwarp_min(mn, 32);
    mx = subwarp_max(mx, 32);
    __shared__ float smn[8], smx[8];
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
        if (mn <= mx) { atomic_min_f32(&state[2 * b], mn); atomic_max_f32(&state[2 * b + 1], mx); }
    }
}
template <int DT>
__global__ void gparam_kernel(const float* __restrict__ state, int64_t batch, float* __restrict__ gs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const float mn = fminf(state[2 * i], 0.0f), mx = fmaxf(state[2 * i + 1], 0.0f);
    gs[i] = gparam<DT>(fmaxf(fabsf(mn), fabsf(mx)));
}
int launch_global_scale(int dt, const void* x, int64_t batch, int64_t numel, float* state, int running, float* gs,
                        cudaStream_t st) {
    B200Q_REQUIRE(batch >= 1 && batch <= 65535, "batch out of range");
    B200Q_REQUIRE(((uintptr_t)x & 15) == 0 && (batch == 1 || (numel * (dt == DT_F32 ? 4 : 2)) % 16 == 0),
                  "tensor must be 16-byte aligned");
    if (!running) { init_minmax_state_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, st>>>(state, batch); B200Q_CHECK_LAUNCH(); }
    if (numel > 0) {
        const int64_t per = (numel / 8 + 255) / 256;
        const unsigned gx = (unsigned)max((int64_t)1, min(per, (int64_t)max(1, kNumSMs * 8 / (int)min(batch, (int64_t)64))));
        B200Q_DISPATCH_DT(dt, { minmax_tensor_kernel<DT><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(x, numel, state); });
        B200Q_CHECK_LAUNCH();
    }
    if (gs) {
        B200Q_DISPATCH_DT(dt, { gparam_kernel<DT><<<(unsigned)((batch + 255) / 256), 256, 0, st>>>(state, batch, gs); });
        B200Q_CHECK_LAUNCH();
    }
    return B200Q_OK;
}

// update_fused_layer_weight_global_scales: `span` consecutive matrices (gate/up of one expert) share min(global_scale)
__global__ void span_min_kernel(float* gs, int64_t n_spans, int span) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_spans) return;
    float m = gs[s * span];
    for (int i = 1; i < span; i++) m = fminf(m, gs[s * span + i]);
    for (int i = 0; i < span; i++) gs[s * span + i] = m;
}
int launch_span_min(float* gs, int64_t batch, int span, cudaStream_t st) {
    const int64_t n = batch / span;
    span_min_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gs, n, span);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

int launch_minmax(int dt, const void* w, int64_t batch, int64_t rows, int64_t cols, int strategy, int group, int bh, int bw,
                  void* mn, void* mx, cudaStream_t st) {
    B200Q_REQUIRE(((uintptr_t)w & 15) == 0, "weight pointer must be 16-byte aligned");
    if (batch * rows * cols == 0) return B200Q_OK;
    if (strategy == ST_GROUP || strategy == ST_CHANNEL) {
        const int g = strategy == ST_CHANNEL ? 0 : group;
        if (g) B200Q_REQUIRE((g == 16 || g == 32 || g == 64 || g == 128 || g == 256) && cols % g == 0,
                             "group_size %d unsupported or does not divide %lld columns", g, (long long)cols);
        B200Q_REQUIRE(cols % 8 == 0, "columns must be a multiple of 8");
        const int64_t nrows = batch * rows;
        static const bool fast_group = getenv("B200Q_MINMAX_LEGACY") == nullptr;  // A/B switch
        if (fast_group && dt == DT_BF16 && g != 0 && cols < (1ll << 31) && (nrows + 7) / 8 < (1ll << 31)) {
            switch (g) {
                case 16: launch_minmax_group_bf16<2>(w, nrows, cols, mn, mx, st); break;
                case 32: launch_minmax_group_bf16<4>(w, nrows, cols, mn, mx, st); break;
                case 64: launch_minmax_group_bf16<8>(w, nrows, cols, mn, mx, st); break;
                case 128: launch_minmax_group_bf16<16>(w, nrows, cols, mn, mx, st); break;
                default: launch_minmax_group_bf16<32>(w, nrows, cols, mn, mx, st); break;
            }
            B200Q_CHECK_LAUNCH();
            return B200Q_OK;
        }
        B200Q_DISPATCH_DT(dt, { minmax_rows_kernel<DT><<<(unsigned)((nrows + 7) / 8), 256, 0, st>>>(w, nrows, cols, g, mn, mx); });
    } else if (strategy == ST_BLOCK) {
        B200Q_REQUIRE(bw % 8 == 0 && cols % 8 == 0 && bh > 0, "block width and columns must be multiples of 8");
        dim3 grid((unsigned)((cols + bw - 1) / bw), (unsigned)((rows + bh - 1) / bh), (unsigned)batch);
        static const bool fast_block = getenv("B200Q_MINMAX_LEGACY") == nullptr;  // A/B switch
        if (fast_block && dt == DT_BF16 && bh == 128 && bw == 128 && grid.y <= 65535 && grid.z <= 65535)
            minmax_block128_bf16_kernel<<<grid, 256, 0, st>>>((const uint16_t*)w, rows, cols, (uint16_t*)mn, (uint16_t*)mx);
        else
            B200Q_DISPATCH_DT(dt, { minmax_block_kernel<DT><<<grid, 256, 0, st>>>(w, rows, cols, bh, bw, mn, mx); });
    } else {
        set_error("minmax: TENSOR strategy goes through b200q_global_scale / the tensor compress workspace");
        return B200Q_EINVAL;
    }
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// ---- O4 MSE observer  (LLMC observers/mse.py, restated in SURVEY.md Appendix A)
// Per quantization chunk (GROUP: g consecutive elements of a row, CHANNEL: a row): for i < n_steps, p = 1 - i / grid:
//     (s, z) = calculate_qparams(T(p * min), T(p * max));  q = fake_quantize(x, s, z)  [strategy patched to TOKEN: one qparam per chunk]
//     err    = T(sum over the chunk of T(|T(q - x)| ^ norm))          -- the reference's in-place chain on tensors of dtype T
//     keep (T(p * min), T(p * max)) where err < best
// and a TENSOR-wide early stop after `patience` consecutive steps without an improvement in ANY chunk.  The early stop only cuts
// the tail of every chunk's search, so one pass computes each chunk's "improved at step i" bit mask for all steps plus the OR of
// the masks over the tensor; the select kernel derives the stop index from the OR and picks each chunk's last improvement before
// it.  The weight is read from HBM once; the search itself is ALU/MUFU work (n_steps fake-quantize + pow per element).
template <int DT, int QT>
__global__ void __launch_bounds__(256) mse_search_kernel(const void* __restrict__ w, int64_t n_chunks, int64_t chunks_per_batch, int len,
                                                         int lanes, int nbits, int symmetric, const float* __restrict__ gs_ptr,
                                                         int n_steps, int grid_n, float norm, void* __restrict__ mn_out,
                                                         void* __restrict__ mx_out, uint32_t* __restrict__ mask_out,
                                                         uint32_t* __restrict__ gmask) {
    const int lane = threadIdx.x & 31;
    const int sub = lane / lanes, sl = lane % lanes;            // sub-warp and lane within it
    const int per_warp = 32 / lanes;
    const int64_t chunk = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * per_warp + sub;
    const bool live = chunk < n_chunks;
    const int64_t base = (live ? chunk : 0) * (int64_t)len;
    const float lo = (QT == QT_INT) ? -(float)(1 << (nbits - 1)) : (QT == QT_FP8 ? -448.0f : -6.0f);
    const float hi = (QT == QT_INT) ? (float)((1 << (nbits - 1)) - 1) : (QT == QT_FP8 ? 448.0f : 6.0f);
    const float gs = gs_ptr ? gs_ptr[0] : 1.0f;
    // Tensor.pow_(python float) casts the exponent to the tensor dtype first (ATen PowKernel: exp_scalar.to<scalar_t>()): a bf16
    // tensor is raised to 2.40625, an fp16 one to 2.400390625
    const float nrm = round_to<DT>(norm);
    // raw min / max of the chunk
    float mn = INFINITY, mx = -INFINITY;
    for (int e = sl; e < len; e += lanes) {
        const float x = load_T<DT>(w, base + e);
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    mn = subwarp_min(mn, lanes);
    mx = subwarp_max(mx, lanes);
    uint32_t mask = 0;
    float best = DT == DT_F16 ? 65504.0f : (DT == DT_BF16 ? 3.3895313892515355e38f : 3.4028234663852886e38f);  // finfo(T).max
    for (int i = 0; i < n_steps; i++) {
        const float p = (float)(1.0 - (double)i / (double)grid_n);
        const float pmn = round_to<DT>(fmul(p, mn)), pmx = round_to<DT>(fmul(p, mx));
        const float cmn = fminf(pmn, 0.0f), cmx = fmaxf(pmx, 0.0f);
        float s, z = 0.0f;
        if (QT == QT_FP4) {
            qparams_fp4<DT>(fmaxf(fabsf(cmn), fabsf(cmx)), gs, s);  // s = s_eff
        } else if (symmetric) {
            s = scale_sym<DT>(fmaxf(fabsf(cmn), fabsf(cmx)), (hi - lo) * 0.5f);
        } else {
            qparams_asym<DT>(cmn, cmx, lo, hi, s, z);
        }
        float err = 0.0f;
        // bf16 fast path (same idea as fastmath.cuh): T(x / s) through the bracketed reciprocal and T(d ^ norm) through bracketed
        // lg2 / ex2 -- when both bracket ends round to the same bf16 the exact IEEE division / powf would too (rounding is
        // monotone); the rare ambiguous element takes the exact path.  ~6x fewer instructions per (element, grid point).
        constexpr bool FAST = DT == DT_BF16 && QT != QT_FP4;
        float r_lo = 0.0f, r_hi = 0.0f;
        bool safe = false;
        if (FAST) {
            const float r = fast::rcp_approx(s);
            r_lo = __fmul_rn(r, 0.99999952316284179688f);
            r_hi = __fmul_rn(r, 1.00000047683715820312f);
            safe = fast::scale_is_safe(__float_as_uint(s));
        }
        for (int e = sl; e < len; e += lanes) {
            const float x = load_T<DT>(w, base + e);  // L1-resident after the min/max pass
            float q;
            if (FAST) {
                const float ua = round_to<DT>(__fmul_rn(x, r_lo)), ub = round_to<DT>(__fmul_rn(x, r_hi));
                float u = (safe && ua == ub) ? ua : round_to<DT>(fdiv(x, s));
                if (QT == QT_INT) {  // quant_int_f + dequant_val with the quotient already rounded to T
                    u = round_to<DT>(fadd(u, z));
                    if (u == u) u = fminf(fmaxf(u, lo), hi);
                    q = dequant_val<DT>(frint(u), s, z, true);
                } else {             // quant_fp8 + dequant_val
                    u = fadd(u, 0.0f);
                    if (u == u) u = fminf(fmaxf(u, -448.0f), 448.0f);
                    q = dequant_val<DT>(e4m3_decode(e4m3_encode(u)), s, 0.0f, true);
                }
            } else if (QT == QT_INT) q = fq_int<DT>(x, s, z, true, lo, hi);
            else if (QT == QT_FP8) q = fq_fp8<DT>(x, s, true);
            else q = fq_fp4<DT>(x, s);
            const float d = fabsf(round_to<DT>(fadd(q, -x)));
            float pw;
            if (FAST) {
                float l;
                asm("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(d));
                const float t = __fmul_rn(nrm, l);
                const float del = __fmaf_rn(fabsf(t), 4.8e-7f, 4.0e-6f);   // lg2 abs error, the product's rounding, ex2's relative error
                float ea, eb;
                asm("ex2.approx.f32 %0, %1;" : "=f"(ea) : "f"(t - del));
                asm("ex2.approx.f32 %0, %1;" : "=f"(eb) : "f"(t + del));
                const float pa = round_to<DT>(ea), pb = round_to<DT>(eb);
                pw = d == 0.0f ? 0.0f : (pa == pb ? pa : round_to<DT>(powf(d, nrm)));
            } else {
                pw = round_to<DT>(powf(d, nrm));
            }
            err = fadd(err, pw);
        }
        for (int o = lanes >> 1; o > 0; o >>= 1) err = fadd(err, __shfl_xor_sync(0xffffffffu, err, o));
        err = round_to<DT>(err);
        if (err < best) { best = err; mask |= 1u << i; }
    }
    if (live && sl == 0) {
        store_T<DT>(mn_out, chunk, mn);
        store_T<DT>(mx_out, chunk, mx);
        mask_out[chunk] = mask;
        atomicOr(&gmask[chunk / chunks_per_batch], mask);
    }
}

template <int DT>
__global__ void __launch_bounds__(256) mse_select_kernel(int64_t n_chunks, int64_t chunks_per_batch, int n_steps, int grid_n, int patience,
                                                         const uint32_t* __restrict__ mask_in, const uint32_t* __restrict__ gmask,
                                                         void* __restrict__ mn_io, void* __restrict__ mx_io) {
    const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= n_chunks) return;
    const uint32_t any = gmask[chunk / chunks_per_batch];
    int executed = n_steps, quiet = 0;
    for (int i = 0; i < n_steps; i++) {
        if (any & (1u << i)) quiet = 0;
        else if (++quiet >= patience) { executed = i + 1; break; }
    }
    const uint32_t sel = mask_in[chunk] & (executed >= 32 ? 0xffffffffu : ((1u << executed) - 1u));
    if (sel == 0) return;  // nothing beat finfo.max: the un-shrunk range stays
    const int i = 31 - __clz(sel);
    const float p = (float)(1.0 - (double)i / (double)grid_n);
    store_T<DT>(mn_io, chunk, round_to<DT>(fmul(p, load_T<DT>(mn_io, chunk))));
    store_T<DT>(mx_io, chunk, round_to<DT>(fmul(p, load_T<DT>(mx_io, chunk))));
}

int launch_mse_minmax(int dt, int qt, int nbits, int symmetric, int strategy, int group, const void* w, int64_t batch, int64_t rows,
                      int64_t cols, const float* gs, float maxshrink, int patience, int grid_n, float norm, void* mn, void* mx,
                      uint32_t* workspace, cudaStream_t st) {
    B200Q_REQUIRE(strategy == ST_GROUP || strategy == ST_CHANNEL, "mse observer: GROUP / TENSOR_GROUP / CHANNEL strategies only");
    B200Q_REQUIRE(grid_n > 0 && patience > 0, "mse observer: grid and patience must be positive");
    const int n_steps = (int)(maxshrink * (float)grid_n);
    B200Q_REQUIRE(n_steps >= 1 && n_steps <= 32, "mse observer: int(maxshrink * grid) must be in [1, 32], got %d", n_steps);
    B200Q_REQUIRE(symmetric || qt == QT_INT, "Asymmetric Quantization is not supported for FP4/FP8 here");
    B200Q_REQUIRE(qt != QT_FP4 || gs, "mse observer: NVFP4 needs the global scale");
    if (batch * rows * cols == 0) return B200Q_OK;
    const int len = strategy == ST_CHANNEL ? (int)cols : group;
    B200Q_REQUIRE(len > 0 && cols % len == 0, "tensor column shape must be divisible by the given group_size %d but got %lld", len, (long long)cols);
    const int64_t chunks_per_batch = rows * (cols / len), n_chunks = batch * chunks_per_batch;
    int lanes = 32;
    while (lanes > 1 && lanes * 16 > len) lanes >>= 1;  // >= 16 elements per lane: the per-grid-point qparams chain is per lane
    const int per_cta = 8 * (32 / lanes);
    uint32_t* gmask = workspace + n_chunks;
    cudaMemsetAsync(gmask, 0, sizeof(uint32_t) * batch, st);
    const unsigned g1 = (unsigned)((n_chunks + per_cta - 1) / per_cta);
    B200Q_DISPATCH_DT(dt, {
        switch (qt) {
        case QT_INT: mse_search_kernel<DT, QT_INT><<<g1, 256, 0, st>>>(w, n_chunks, chunks_per_batch, len, lanes, nbits, symmetric, gs, n_steps, grid_n, norm, mn, mx, workspace, gmask); break;
        case QT_FP8: mse_search_kernel<DT, QT_FP8><<<g1, 256, 0, st>>>(w, n_chunks, chunks_per_batch, len, lanes, nbits, symmetric, gs, n_steps, grid_n, norm, mn, mx, workspace, gmask); break;
        case QT_FP4: mse_search_kernel<DT, QT_FP4><<<g1, 256, 0, st>>>(w, n_chunks, chunks_per_batch, len, lanes, nbits, symmetric, gs, n_steps, grid_n, norm, mn, mx, workspace, gmask); break;
        default: set_error("bad qtype %d", qt); return B200Q_EINVAL;
        }
    });
    B200Q_CHECK_LAUNCH();
    B200Q_DISPATCH_DT(dt, { mse_select_kernel<DT><<<(unsigned)((n_chunks + 255) / 256), 256, 0, st>>>(n_chunks, chunks_per_batch, n_steps, grid_n, patience, workspace, gmask, mn, mx); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// ---- Q1 calculate_qparams
template <int DT, int QT>
__global__ void __launch_bounds__(256) qparams_kernel(const void* __restrict__ mn_in, const void* __restrict__ mx_in, int64_t n,
                                                      int nbits, int symmetric, const float* __restrict__ gs,
                                                      void* __restrict__ scale, int8_t* __restrict__ zp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lo = (QT == QT_INT) ? -(float)(1 << (nbits - 1)) : (QT == QT_FP8 ? -448.0f : -6.0f);
    const float hi = (QT == QT_INT) ? (float)((1 << (nbits - 1)) - 1) : (QT == QT_FP8 ? 448.0f : 6.0f);
    float mn = fminf(load_T<DT>(mn_in, i), 0.0f), mx = fmaxf(load_T<DT>(mx_in, i), 0.0f);
    float s, z = 0.0f;
    if (symmetric) {
        const float a = fmaxf(fabsf(mn), fabsf(mx));
        if (QT == QT_FP4 && gs) {  // e4m3-rounded fp32 scale, helpers.py:101-126
            float se;
            const uint8_t code = qparams_fp4<DT>(a, gs[0], se);
            ((float*)scale)[i] = e4m3_decode(code);
            if (zp) zp[i] = 0;
            return;
        }
        s = round_to<DT>(fdiv(a, (hi - lo) * 0.5f));
        if (gs) {
            float sf = fmul(gs[0], s);
            ((float*)scale)[i] = sf == 0.0f ? 1.1920928955078125e-07f : sf;
            if (zp) zp[i] = 0;
            return;
        }
        if (s == 0.0f) s = eps_of<DT>();
    } else {
        qparams_asym<DT>(mn, mx, lo, hi, s, z);
    }
    store_T<DT>(scale, i, s);
    if (zp) zp[i] = (int8_t)(int)z;
}

int launch_qparams(int dt, int qt, int nbits, int symmetric, const void* mn, const void* mx, int64_t n, const float* gs,
                   void* scale, int8_t* zp, cudaStream_t st) {
    if (n == 0) return B200Q_OK;
    B200Q_REQUIRE(symmetric || qt == QT_INT, "Asymmetric Quantization is not supported for FP4/FP8 here");
    const unsigned g = (unsigned)((n + 255) / 256);
    B200Q_DISPATCH_DT(dt, {
        switch (qt) {
        case QT_INT: qparams_kernel<DT, QT_INT><<<g, 256, 0, st>>>(mn, mx, n, nbits, symmetric, gs, scale, zp); break;
        case QT_FP8: qparams_kernel<DT, QT_FP8><<<g, 256, 0, st>>>(mn, mx, n, nbits, symmetric, gs, scale, zp); break;
        case QT_FP4: qparams_kernel<DT, QT_FP4><<<g, 256, 0, st>>>(mn, mx, n, nbits, symmetric, gs, scale, zp); break;
        default: set_error("bad qtype %d", qt); return B200Q_EINVAL;
        }
    });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace b200q
