// quant_tile.cu -- fused compress kernels whose quantization chunk is larger than a warp segment:
//   CHANNEL (one scale per row)        INT4/INT8 pack-quantized and FP8 float-quantized
//   BLOCK 128x128 (FP8_BLOCK preset)   CT:quantization/lifecycle/forward_helpers.py:57-110
//   TENSOR (one scale per weight)      FP8
// The whole chunk is held in registers between the reduction and the quantization (8 x 128-bit loads in
// flight per thread), so the weight is read from HBM exactly once.
#include "common.cuh"
#include "kernels.cuh"

namespace b200q {

constexpr int NC = 8;  // register-resident chunks per thread

__device__ __forceinline__ float block_max(float v, float* sm) {
    v = subwarp_max(v, 32);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sm[0];
#pragma unroll
    for (int i = 1; i < 8; i++) r = fmaxf(r, sm[i]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_min(float v, float* sm) {
    v = subwarp_min(v, 32);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sm[0];
#pragma unroll
    for (int i = 1; i < 8; i++) r = fminf(r, sm[i]);
    __syncthreads();
    return r;
}

template <int DT, int QT>
__device__ __forceinline__ void emit_codes(const float x[8], float s, float z, bool use_zp, float lo, float hi, int nbits,
                                           void* out, int64_t e0) {
    if (QT == QT_INT) {
        int c[8];
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = quant_int<DT>(x[i], s, z, use_zp, lo, hi);
        if (nbits == 4) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) w |= ((uint32_t)(c[i] + 8) & 0xfu) << (4 * i);
            stg_stream((uint32_t*)out + (e0 >> 3), w);
        } else {
            uint32_t w0 = 0, w1 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                w0 |= ((uint32_t)(c[i] + 128) & 0xffu) << (8 * i);
                w1 |= ((uint32_t)(c[i + 4] + 128) & 0xffu) << (8 * i);
            }
            stg_stream((uint8_t*)out + e0, make_uint2(w0, w1));
        }
    } else {
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            w0 |= (uint32_t)quant_fp8<DT>(x[i], s, use_zp) << (8 * i);
            w1 |= (uint32_t)quant_fp8<DT>(x[i + 4], s, use_zp) << (8 * i);
        }
        stg_stream((uint8_t*)out + e0, make_uint2(w0, w1));
    }
}

// ------------------------------------------------------------------ CHANNEL: one CTA (256 thr) per row
template <int DT, int QT>
__global__ void __launch_bounds__(256) channel_compress_kernel(const TileParams p) {
    __shared__ float sm[8];
    const int64_t row = blockIdx.x;  // over batch*rows
    const int64_t rbase = row * p.cols;
    const float lo = (QT == QT_INT) ? -(float)(1 << (p.nbits - 1)) : -448.0f;
    const float hi = (QT == QT_INT) ? (float)((1 << (p.nbits - 1)) - 1) : 448.0f;
    const int64_t nch = p.cols / 8;
    Chunk8<DT> ch[NC];
    float mn = INFINITY, mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int64_t c = (int64_t)j * 256 + threadIdx.x;
        if (c < nch) {
            load_chunk<DT>(ch[j], p.w, rbase + c * 8);
            float x[8];
            chunk_to_float<DT>(ch[j], x);
#pragma unroll
            for (int i = 0; i < 8; i++) { mn = fminf(mn, x[i]); mx = fmaxf(mx, x[i]); }
        }
    }
    for (int64_t c = (int64_t)NC * 256 + threadIdx.x; c < nch; c += 256) {  // rows longer than 16384: stream the rest
        Chunk8<DT> t;
        float x[8];
        load_chunk<DT>(t, p.w, rbase + c * 8);
        chunk_to_float<DT>(t, x);
#pragma unroll
        for (int i = 0; i < 8; i++) { mn = fminf(mn, x[i]); mx = fmaxf(mx, x[i]); }
    }
    mn = block_min(mn, sm);
    mx = block_max(mx, sm);
    float s, z = 0.0f;
    const bool asym = (QT == QT_INT) && !p.symmetric;
    if (asym) qparams_asym<DT>(mn, mx, lo, hi, s, z);
    else s = scale_sym<DT>(fmaxf(fabsf(fminf(mn, 0.0f)), fabsf(fmaxf(mx, 0.0f))), QT == QT_INT ? (hi - lo) * 0.5f : hi);
    const bool use_zp = (QT == QT_INT) ? asym : (p.has_zp != 0);
    if (threadIdx.x == 0) {
        store_T<DT>(p.scale, row, s);
        if (asym) {  // pack_to_int32(zero_point, packed_dim=0): nibble (row % pf) of word (row / pf) within this matrix
            const int pf = 32 / p.nbits;
            const int64_t b = row / p.rows, r = row % p.rows;
            const int64_t zrows = (p.rows + pf - 1) / pf;
            atomicOr((unsigned int*)&p.zp_packed[b * zrows + r / pf],
                     ((uint32_t)((int)z + (1 << (p.nbits - 1))) & ((1u << p.nbits) - 1u)) << (p.nbits * (int)(r % pf)));
        }
    }
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int64_t c = (int64_t)j * 256 + threadIdx.x;
        if (c < nch) {
            float x[8];
            chunk_to_float<DT>(ch[j], x);
            emit_codes<DT, QT>(x, s, z, use_zp, lo, hi, p.nbits, p.out, rbase + c * 8);
        }
    }
    for (int64_t c = (int64_t)NC * 256 + threadIdx.x; c < nch; c += 256) {
        Chunk8<DT> t;
        float x[8];
        load_chunk<DT>(t, p.w, rbase + c * 8);
        chunk_to_float<DT>(t, x);
        emit_codes<DT, QT>(x, s, z, use_zp, lo, hi, p.nbits, p.out, rbase + c * 8);
    }
}

int launch_channel_compress(int dt, int qt, const TileParams& p, int64_t batch, cudaStream_t st) {
    B200Q_REQUIRE(qt == QT_INT || qt == QT_FP8, "channel compress supports INT and FP8");
    B200Q_REQUIRE(p.cols % 8 == 0, "columns must be a multiple of 8, got %lld", (long long)p.cols);
    B200Q_REQUIRE(((uintptr_t)p.w & 15) == 0, "weight pointer must be 16-byte aligned");
    if (batch * p.rows * p.cols == 0) return B200Q_OK;
    if (qt == QT_INT && !p.symmetric) {
        const int pf = 32 / p.nbits;
        B200Q_REQUIRE(p.zp_packed != nullptr, "Asymmetric quant requires zero-point values");
        cudaMemsetAsync(p.zp_packed, 0, sizeof(int32_t) * batch * ((p.rows + pf - 1) / pf), st);
    }
    const unsigned grid = (unsigned)(batch * p.rows);
    B200Q_DISPATCH_DT(dt, {
        if (qt == QT_INT) channel_compress_kernel<DT, QT_INT><<<grid, 256, 0, st>>>(p);
        else channel_compress_kernel<DT, QT_FP8><<<grid, 256, 0, st>>>(p);
    });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// ------------------------------------------------------------------ BLOCK 128x128 FP8: one CTA per tile
// thread t, step i: tile row (i*256+t)/16, 8-column chunk (i*256+t)%16 -> a warp reads 2 rows x 256 B.
template <int DT>
__global__ void __launch_bounds__(256) block_fp8_kernel(const TileParams p) {
    __shared__ float sm[8];
    const int64_t b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * 128, c0 = (int64_t)blockIdx.x * 128;
    const int64_t mat = b * p.rows * p.cols;
    Chunk8<DT> ch[NC];
    float a = 0.0f;
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const int id = i * 256 + threadIdx.x;
        const int64_t r = r0 + (id >> 4), c = c0 + (int64_t)(id & 15) * 8;
        if (r < p.rows && c < p.cols) {
            load_chunk<DT>(ch[i], p.w, mat + r * p.cols + c);
            float x[8];
            chunk_to_float<DT>(ch[i], x);
#pragma unroll
            for (int k = 0; k < 8; k++) a = fmaxf(a, fabsf(x[k]));
        }
    }
    a = block_max(a, sm);  // zero padding of ragged edges never raises |.|max
    const float s = scale_sym<DT>(a, 448.0f);
    if (threadIdx.x == 0) store_T<DT>(p.scale, (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x, s);
    const bool use_zp = p.has_zp != 0;
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const int id = i * 256 + threadIdx.x;
        const int64_t r = r0 + (id >> 4), c = c0 + (int64_t)(id & 15) * 8;
        if (r < p.rows && c < p.cols) {
            float x[8];
            chunk_to_float<DT>(ch[i], x);
            emit_codes<DT, QT_FP8>(x, s, 0.0f, use_zp, -448.0f, 448.0f, 8, p.out, mat + r * p.cols + c);
        }
    }
}

int launch_block_fp8_compress(int dt, const TileParams& p, int64_t batch, cudaStream_t st) {
    B200Q_REQUIRE(p.cols % 8 == 0, "columns must be a multiple of 8, got %lld", (long long)p.cols);
    B200Q_REQUIRE(((uintptr_t)p.w & 15) == 0, "weight pointer must be 16-byte aligned");
    if (batch * p.rows * p.cols == 0) return B200Q_OK;
    dim3 grid((unsigned)((p.cols + 127) / 128), (unsigned)((p.rows + 127) / 128), (unsigned)batch);
    B200Q_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "too many row blocks / batch entries for one launch");
    B200Q_DISPATCH_DT(dt, { block_fp8_kernel<DT><<<grid, 256, 0, st>>>(p); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

// ------------------------------------------------------------------ TENSOR FP8: absmax pass + quantize pass
__global__ void zero_f32_kernel(float* p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0f;
}
template <int DT>
__global__ void __launch_bounds__(256) tensor_absmax_kernel(const void* __restrict__ w, int64_t numel, float* __restrict__ ws) {
    __shared__ float sm[8];
    const int64_t b = blockIdx.y;
    float a = 0.0f;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < numel / 8; t += (int64_t)gridDim.x * blockDim.x) {
        Chunk8<DT> ch;
        float x[8];
        load_chunk<DT>(ch, w, b * numel + t * 8);
        chunk_to_float<DT>(ch, x);
#pragma unroll
        for (int k = 0; k < 8; k++) a = fmaxf(a, fabsf(x[k]));
    }
    a = block_max(a, sm);
    if (threadIdx.x == 0) atomicMax((int*)&ws[b], __float_as_int(a));  // a >= 0: int order == float order
}
template <int DT>
__global__ void __launch_bounds__(256) tensor_fp8_quant_kernel(const TileParams p, int64_t numel) {
    const int64_t b = blockIdx.y;
    const float s = scale_sym<DT>(p.workspace[b], 448.0f);
    if (blockIdx.x == 0 && threadIdx.x == 0) store_T<DT>(p.scale, b, s);
    const bool use_zp = p.has_zp != 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < numel / 8; t += (int64_t)gridDim.x * blockDim.x) {
        Chunk8<DT> ch;
        float x[8];
        load_chunk<DT>(ch, p.w, b * numel + t * 8);
        chunk_to_float<DT>(ch, x);
        emit_codes<DT, QT_FP8>(x, s, 0.0f, use_zp, -448.0f, 448.0f, 8, p.out, b * numel + t * 8);
    }
}

int launch_tensor_fp8_compress(int dt, const TileParams& p, int64_t batch, cudaStream_t st) {
    const int64_t numel = p.rows * p.cols;
    B200Q_REQUIRE(numel % 8 == 0, "rows*cols must be a multiple of 8");
    B200Q_REQUIRE(p.workspace != nullptr, "TENSOR strategy needs a workspace of 4*batch bytes");
    B200Q_REQUIRE(batch <= 65535, "batch out of range");
    if (batch * numel == 0) return B200Q_OK;
    zero_f32_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, st>>>(p.workspace, batch);
    const unsigned gx = (unsigned)max((int64_t)1, min((numel / 8 + 255) / 256, (int64_t)kNumSMs * 8));
    B200Q_DISPATCH_DT(dt, {
        tensor_absmax_kernel<DT><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(p.w, numel, p.workspace);
        tensor_fp8_quant_kernel<DT><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(p, numel);
    });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace b200q
