// pack.cu -- pack_to_int32 / unpack_from_int32 (CT:compressors/pack_quantized/helpers.py:20-161) and
// pack_fp4_to_uint8 / unpack_fp4_from_uint8 (CT:compressors/nvfp4/helpers.py:34-111) as stand-alone kernels.
// The fused compress kernels never call these; they exist for the un-fused CT call sites and decompression.
#include "common.cuh"
#include "kernels.cuh"

namespace b200q {

// one thread per output word; packed_dim 1: words along columns, 0: along rows (used for zero-points)
__global__ void __launch_bounds__(256) pack_int32_kernel(const int8_t* __restrict__ v, int64_t rows, int64_t cols, int nbits,
                                                         int packed_dim, int32_t* __restrict__ out) {
    const int pf = 32 / nbits, off = 1 << (nbits - 1);
    const int64_t orow = packed_dim == 1 ? rows : (rows + pf - 1) / pf;
    const int64_t ocol = packed_dim == 1 ? (cols + pf - 1) / pf : cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < orow * ocol; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / ocol, c = t % ocol;
        uint32_t acc = 0;
        for (int i = 0; i < pf; i++) {
            const int64_t rr = packed_dim == 1 ? r : r * pf + i, cc = packed_dim == 1 ? c * pf + i : c;
            if (rr < rows && cc < cols) acc += ((uint32_t)(uint8_t)(v[rr * cols + cc] + off)) << (nbits * i);  // CT sums the shifted bytes (helpers.py:20-90): codes outside the range carry
        }
        out[t] = (int32_t)acc;
    }
}
__global__ void __launch_bounds__(256) unpack_int32_kernel(const int32_t* __restrict__ p, int64_t rows, int64_t cols, int nbits,
                                                           int packed_dim, int8_t* __restrict__ out) {
    const int pf = 32 / nbits, off = 1 << (nbits - 1);
    const uint32_t mask = (1u << nbits) - 1u;
    const int64_t pcols = packed_dim == 1 ? (cols + pf - 1) / pf : cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows * cols; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / cols, c = t % cols;
        const uint32_t w = packed_dim == 1 ? (uint32_t)p[r * pcols + c / pf] : (uint32_t)p[(r / pf) * pcols + c];
        const int sh = nbits * (int)(packed_dim == 1 ? c % pf : r % pf);
        out[t] = (int8_t)((int)((w >> sh) & mask) - off);
    }
}

// nearest e2m1 magnitude (first minimum, like torch.argmin over |abs(x) - table| evaluated in T) | signbit << 3
template <int DT>
__global__ void __launch_bounds__(256) pack_fp4_kernel(const void* __restrict__ x, int64_t n_pairs, uint8_t* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += (int64_t)gridDim.x * blockDim.x) {
        uint32_t nib[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const float v = load_T<DT>(x, 2 * t + k);
            const float a = fabsf(v);
            int best = 0;
            float bd = INFINITY;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float d = fabsf(round_to<DT>(fadd(a, -e2m1_value(i))));
                if (d < bd) { bd = d; best = i; }
            }
            nib[k] = (uint32_t)best | ((__float_as_uint(v) >> 31) << 3);
        }
        out[t] = (uint8_t)(nib[0] | (nib[1] << 4));
    }
}
template <int DT>
__global__ void __launch_bounds__(256) unpack_fp4_kernel(const uint8_t* __restrict__ p, int64_t n, void* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t b = p[t >> 1];
        const uint32_t nib = (t & 1) ? (b >> 4) : (b & 0xf);
        const float v = e2m1_value(nib & 7u);
        store_T<DT>(out, t, (nib & 8u) ? -v : v);
    }
}

// ---- fused decompress: packed codes + qparams -> T weights in one pass (CT Compressor.decompress =
// unpack_from_int32 / unpack_fp4_from_uint8 followed by dequantize, compressors/pack_quantized/base.py:79-113, nvfp4/base.py:74-96).
// One thread per output chunk of 8 elements (16-byte store); reads 0.5 B + qparams per element instead of the int8 / T
// intermediates of the two-step path (1 B resp. 2 B written and read back per element).
template <int DT>
__global__ void __launch_bounds__(256) decompress_int_packed_kernel(const int32_t* __restrict__ packed, const void* __restrict__ scale,
                                                                    const int32_t* __restrict__ zp_packed, int64_t batch, int64_t rows,
                                                                    int64_t cols, int group, int nbits, void* __restrict__ out) {
    const int pf = 32 / nbits, off = 1 << (nbits - 1);
    const uint32_t mask = (1u << nbits) - 1u;
    const int64_t pcols = (cols + pf - 1) / pf, chunks_per_row = (cols + 7) / 8;
    const int64_t gtot = group > 0 ? cols / group : 1, zrows = (rows + pf - 1) / pf;
    const int64_t total = batch * rows * chunks_per_row;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t br = t / chunks_per_row, ch = t - br * chunks_per_row;  // br = b * rows + r
        const int64_t b = br / rows, r = br - b * rows, c0 = ch * 8;
        float y[8];
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int64_t c = c0 + i;
            if (c >= cols) { y[i] = 0.0f; continue; }
            if (i % pf == 0 || i == 0) w = (uint32_t)packed[br * pcols + c / pf];
            const float q = (float)((int)((w >> (nbits * (int)(c % pf))) & mask) - off);
            const int64_t g = group > 0 ? c / group : 0;
            const float s = load_T<DT>(scale, br * gtot + g);
            float z = 0.0f;
            if (zp_packed) {
                const uint32_t zw = (uint32_t)zp_packed[(b * zrows + r / pf) * gtot + g];
                z = (float)((int)((zw >> (nbits * (int)(r % pf))) & mask) - off);
            }
            y[i] = dequant_val<DT>(q, s, z, zp_packed != nullptr);
        }
        if (c0 + 8 <= cols) store_chunk_T<DT>(out, br * cols + c0, y);
        else for (int i = 0; c0 + i < cols; i++) store_T<DT>(out, br * cols + c0 + i, y[i]);
    }
}
// NVFP4: one thread per group of 16 (8 bytes of codes, one e4m3 scale): value * (fp32(scale) / gs), fp32 product cast to T
template <int DT>
__global__ void __launch_bounds__(256) decompress_nvfp4_kernel(const uint2* __restrict__ packed, const uint8_t* __restrict__ scale,
                                                               const float* __restrict__ gs, int gs_stride, int64_t groups_per_mat,
                                                               int64_t n_groups, void* __restrict__ out) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const uint2 cw = packed[g];
        // the module holds the e4m3 scale as T (exact) and dequantize() promotes scale / global_scale to fp32
        const float se = fdiv(round_to<DT>(e4m3_decode(scale[g])), gs[gs_stride ? g / groups_per_mat : 0]);
        const uint32_t words[2] = {cw.x, cw.y};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t nib = (words[h] >> (4 * i)) & 0xfu;
                const float v = e2m1_value(nib & 7u);
                y[i] = round_to<DT>(fmul((nib & 8u) ? -v : v, se));
            }
            store_chunk_T<DT>(out, g * 16 + h * 8, y);
        }
    }
}

// ------------------------------------------------------------------------------------------------ flat fast paths
// 4-bit codes along the columns with cols % 8 == 0 (resp. bf16 NVFP4 values with cols % 8 == 0): both sides of the conversion are
// contiguous, so the tensor is a flat array of 8-element chunks.  One thread converts U chunks per step with every load issued first
// (the per-element kernels above index with 64-bit divisions: 0.06-0.32 of the HBM roofline).  Inputs outside the format's domain
// (int8 codes outside [-8, 7], values that are not e2m1 grid points) are redone with the element-wise arithmetic of the kernels above.
constexpr int PK_U = 4;

__device__ __forceinline__ uint32_t pk_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// 16 int8 codes (uint4) -> 2 packed words
__global__ void __launch_bounds__(256) pack4_flat_kernel(const uint4* __restrict__ v, int64_t n16, uint2* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < n16; t0 += stride * PK_U) {
        uint4 x[PK_U];
#pragma unroll
        for (int u = 0; u < PK_U; u++) x[u] = t0 + u * stride < n16 ? ldg_stream(v + t0 + u * stride) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < PK_U; u++) {
            const int64_t t = t0 + u * stride;
            if (t >= n16) continue;
            const uint32_t w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
            uint32_t b[4], bad = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                b[k] = ((w[k] & 0x7f7f7f7fu) + 0x08080808u) ^ (w[k] & 0x80808080u);  // per-byte (code + 8) mod 256
                bad |= b[k] & 0xf0f0f0f0u;
            }
            uint32_t o[2];
            if (bad == 0) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t lo = b[2 * h] | (b[2 * h] >> 4), hi = b[2 * h + 1] | (b[2 * h + 1] >> 4);  // bytes 0, 2: two nibbles each
                    o[h] = pk_prmt(lo, hi, 0x6420);
                }
            } else {  // the reference SUMS (code + 8) << 4 i without masking
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) acc += ((b[2 * h + (i >> 2)] >> (8 * (i & 3))) & 0xffu) << (4 * i);
                    o[h] = acc;
                }
            }
            stg_stream(out + t, make_uint2(o[0], o[1]));
        }
    }
}
// 4 packed words (uint4) -> 32 int8 codes
__global__ void __launch_bounds__(256) unpack4_flat_kernel(const uint4* __restrict__ p, int64_t n4, uint4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < n4; t0 += stride * PK_U) {
        uint4 x[PK_U];
#pragma unroll
        for (int u = 0; u < PK_U; u++) x[u] = t0 + u * stride < n4 ? __ldg(p + t0 + u * stride) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < PK_U; u++) {
            const int64_t t = t0 + u * stride;
            if (t >= n4) continue;
            const uint32_t w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t lo = w[k] & 0x0f0f0f0fu, hi = (w[k] >> 4) & 0x0f0f0f0fu;
                // nibble - 8 as int8: two's-complement nibble (xor 8), bit 3 * 0x1e fills the high nibble
                const uint32_t a = pk_prmt(lo, hi, 0x5140) ^ 0x08080808u, c = pk_prmt(lo, hi, 0x7362) ^ 0x08080808u;
                o[2 * k] = a | ((a & 0x08080808u) * 0x1eu);
                o[2 * k + 1] = c | ((c & 0x08080808u) * 0x1eu);
            }
            stg_stream(out + 2 * t, make_uint4(o[0], o[1], o[2], o[3]));
            stg_stream(out + 2 * t + 1, make_uint4(o[4], o[5], o[6], o[7]));
        }
    }
}
// 4 bytes of e2m1 code pairs -> 8 bf16 values
__global__ void __launch_bounds__(256) unpack_fp4_flat_kernel(const uint32_t* __restrict__ p, int64_t nw, uint4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < nw; t0 += stride * PK_U) {
        uint32_t x[PK_U];
#pragma unroll
        for (int u = 0; u < PK_U; u++) x[u] = t0 + u * stride < nw ? __ldg(p + t0 + u * stride) : 0u;
#pragma unroll
        for (int u = 0; u < PK_U; u++) {
            const int64_t t = t0 + u * stride;
            if (t >= nw) continue;
            uint32_t y[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float lo, hi;
                asm("{ .reg .b8 t; .reg .b16 u, l, h; .reg .b32 r; cvt.u16.u32 u, %2; cvt.u8.u16 t, u; cvt.rn.f16x2.e2m1x2 r, t; mov.b32 {l, h}, r; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }"
                    : "=f"(lo), "=f"(hi) : "r"((x[u] >> (8 * k)) & 0xffu));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y[k]) : "f"(hi), "f"(lo));
            }
            stg_stream(out + t, make_uint4(y[0], y[1], y[2], y[3]));
        }
    }
}
// 8 bf16 values (uint4) -> 4 bytes of e2m1 code pairs.  A value that IS a grid point (the output of CT quantize) converts exactly with
// cvt.rn.satfinite.e2m1x2; anything else takes the reference's first-minimum search in bf16.
__device__ __noinline__ uint32_t pack_fp4_exact8(const uint4 raw) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t o = 0;
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        const float v = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
        const float a = fabsf(v);
        int best = 0;
        float bd = INFINITY;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float d = fabsf(round_to<DT_BF16>(fadd(a, -e2m1_value(i))));
            if (d < bd) { bd = d; best = i; }
        }
        o |= ((uint32_t)best | ((__float_as_uint(v) >> 31) << 3)) << (4 * e);
    }
    return o;
}
__global__ void __launch_bounds__(256) pack_fp4_flat_kernel(const uint4* __restrict__ v, int64_t n8, uint32_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < n8; t0 += stride * PK_U) {
        uint4 x[PK_U];
#pragma unroll
        for (int u = 0; u < PK_U; u++) x[u] = t0 + u * stride < n8 ? ldg_stream(v + t0 + u * stride) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < PK_U; u++) {
            const int64_t t = t0 + u * stride;
            if (t >= n8) continue;
            const uint32_t w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
            uint32_t o = 0;
            bool on_grid = true;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float lo = __uint_as_float(w[k] << 16), hi = __uint_as_float(w[k] & 0xffff0000u);
                uint32_t code;
                asm("{ .reg .b8 t; .reg .b16 u; cvt.rn.satfinite.e2m1x2.f32 t, %1, %2; cvt.u16.u8 u, t; cvt.u32.u16 %0, u; }" : "=r"(code) : "f"(hi), "f"(lo));
                float dl, dh;
                asm("{ .reg .b8 t; .reg .b16 u, l, h; .reg .b32 r; cvt.u16.u32 u, %2; cvt.u8.u16 t, u; cvt.rn.f16x2.e2m1x2 r, t; mov.b32 {l, h}, r; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }"
                    : "=f"(dl), "=f"(dh) : "r"(code));
                on_grid = on_grid && __float_as_uint(dl) == __float_as_uint(lo) && __float_as_uint(dh) == __float_as_uint(hi);
                o |= code << (8 * k);
            }
            if (!on_grid) o = pack_fp4_exact8(x[u]);
            stg_stream(out + t, o);
        }
    }
}
static unsigned flat_grid(int64_t work_items) {  // threads each take PK_U items per step; at most 4 CTAs per SM in one wave
    return (unsigned)max((int64_t)1, min((int64_t)kNumSMs * 4, (work_items + 256 * PK_U - 1) / (256 * PK_U)));
}

static int nblocks(int64_t work) { return (int)max((int64_t)1, min((int64_t)kNumSMs * 16, (work + 255) / 256)); }

int launch_pack_int32(const int8_t* v, int64_t rows, int64_t cols, int nbits, int packed_dim, int32_t* out, cudaStream_t st) {
    B200Q_REQUIRE(nbits >= 1 && nbits <= 8, "Packing is only supported for less than 8 bits");
    B200Q_REQUIRE(packed_dim == 0 || packed_dim == 1, "packed_dim must be 0 or 1");
    if (rows * cols == 0) return B200Q_OK;
    if (nbits == 4 && packed_dim == 1 && cols % 8 == 0 && (rows * cols) % 16 == 0 && (((uintptr_t)v) & 15) == 0 && (((uintptr_t)out) & 7) == 0) {
        pack4_flat_kernel<<<flat_grid(rows * cols / 16), 256, 0, st>>>((const uint4*)v, rows * cols / 16, (uint2*)out);
        B200Q_CHECK_LAUNCH();
        return B200Q_OK;
    }
    pack_int32_kernel<<<nblocks(rows * cols / (32 / nbits) + 1), 256, 0, st>>>(v, rows, cols, nbits, packed_dim, out);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}
int launch_unpack_int32(const int32_t* p, int64_t rows, int64_t cols, int nbits, int packed_dim, int8_t* out, cudaStream_t st) {
    B200Q_REQUIRE(nbits >= 1 && nbits <= 8, "Unpacking is only supported for less than 8 bits");
    if (rows * cols == 0) return B200Q_OK;
    if (nbits == 4 && packed_dim == 1 && cols % 8 == 0 && (rows * cols) % 32 == 0 && (((uintptr_t)p) & 15) == 0 && (((uintptr_t)out) & 15) == 0) {
        unpack4_flat_kernel<<<flat_grid(rows * cols / 32), 256, 0, st>>>((const uint4*)p, rows * cols / 32, (uint4*)out);
        B200Q_CHECK_LAUNCH();
        return B200Q_OK;
    }
    unpack_int32_kernel<<<nblocks(rows * cols), 256, 0, st>>>(p, rows, cols, nbits, packed_dim, out);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}
int launch_pack_fp4(int dt, const void* x, int64_t rows, int64_t cols, uint8_t* out, cudaStream_t st) {
    B200Q_REQUIRE(cols % 2 == 0, "tensor must have an even number of columns for nvfp4 compression");
    if (rows * cols == 0) return B200Q_OK;
    if (dt == DT_BF16 && (rows * cols) % 8 == 0 && (((uintptr_t)x) & 15) == 0 && (((uintptr_t)out) & 3) == 0) {
        pack_fp4_flat_kernel<<<flat_grid(rows * cols / 8), 256, 0, st>>>((const uint4*)x, rows * cols / 8, (uint32_t*)out);
        B200Q_CHECK_LAUNCH();
        return B200Q_OK;
    }
    B200Q_DISPATCH_DT(dt, { pack_fp4_kernel<DT><<<nblocks(rows * cols / 2), 256, 0, st>>>(x, rows * cols / 2, out); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}
int launch_unpack_fp4(int dt, const uint8_t* p, int64_t rows, int64_t cols, void* out, cudaStream_t st) {
    if (rows * cols == 0) return B200Q_OK;
    if (dt == DT_BF16 && (rows * cols) % 8 == 0 && (((uintptr_t)p) & 3) == 0 && (((uintptr_t)out) & 15) == 0) {
        unpack_fp4_flat_kernel<<<flat_grid(rows * cols / 8), 256, 0, st>>>((const uint32_t*)p, rows * cols / 8, (uint4*)out);
        B200Q_CHECK_LAUNCH();
        return B200Q_OK;
    }
    B200Q_DISPATCH_DT(dt, { unpack_fp4_kernel<DT><<<nblocks(rows * cols), 256, 0, st>>>(p, rows * cols, out); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

int launch_decompress_int_packed(int dt, const int32_t* packed, const void* scale, const int32_t* zp_packed, int64_t batch, int64_t rows,
                                 int64_t cols, int group, int nbits, void* out, cudaStream_t st) {
    B200Q_REQUIRE(nbits == 4 || nbits == 8, "pack-quantized decompress supports 4 and 8 bits");
    B200Q_REQUIRE(group == 0 || cols % group == 0, "tensor column shape must be divisible by the given group_size %d but got %lld", group,
                  (long long)cols);
    B200Q_REQUIRE(cols % 8 == 0 || (dt == DT_F32), "columns must be a multiple of 8 for the 16-byte stores");
    if (batch * rows * cols == 0) return B200Q_OK;
    B200Q_DISPATCH_DT(dt, { decompress_int_packed_kernel<DT><<<nblocks(batch * rows * ((cols + 7) / 8)), 256, 0, st>>>(packed, scale, zp_packed, batch,
                                                                                                                rows, cols, group, nbits, out); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}
int launch_decompress_nvfp4(int dt, const uint8_t* packed, const uint8_t* scale, const float* gs, int gs_stride, int64_t batch, int64_t rows,
                            int64_t cols, void* out, cudaStream_t st) {
    B200Q_REQUIRE(cols % 16 == 0, "tensor column shape must be divisible by the given group_size 16 but got %lld", (long long)cols);
    B200Q_REQUIRE((((uintptr_t)packed) & 7) == 0 && (((uintptr_t)out) & 15) == 0, "packed / out must be 8 / 16-byte aligned");
    if (batch * rows * cols == 0) return B200Q_OK;
    const int64_t gpm = rows * (cols / 16), n = batch * gpm;
    B200Q_DISPATCH_DT(dt, { decompress_nvfp4_kernel<DT><<<nblocks(n), 256, 0, st>>>((const uint2*)packed, scale, gs, gs_stride, gpm, n, out); });
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace b200q
