// common.cuh -- device helpers shared by the streaming kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "qmath.cuh"

namespace b200q {

// ---- error plumbing for the C ABI (thread-local message, errno-style negative codes)
void set_error(const char* fmt, ...);
#define B200Q_OK 0
#define B200Q_EINVAL (-22)
#define B200Q_ECUDA (-5)
#define B200Q_ENOSYS (-38)

#define B200Q_CHECK_LAUNCH()                                                     \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            set_error("%s:%d CUDA launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return B200Q_ECUDA;                                                  \
        }                                                                        \
    } while (0)

#define B200Q_REQUIRE(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            set_error(__VA_ARGS__);              \
            return B200Q_EINVAL;                 \
        }                                        \
    } while (0)

constexpr int kNumSMs = 148;  // B200

template <int DT> struct ElemSize { static constexpr int value = (DT == DT_F32) ? 4 : 2; };

// ---- streaming 128-bit global accesses (read-once data: do not allocate in L1)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// read-twice data (first pass of a two-pass kernel): keep the line in the L2 for the second pass -- the streaming load above tags
// its lines evict-first, which made the L2 drop them before the re-read (ncu: 31 % L2 read hit rate on the fused NVFP4 kernel)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_keep(const void* p, uint64_t policy) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ void stg_stream(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y));
}
__device__ __forceinline__ void stg_stream(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v));
}

// ---- an 8-element chunk of T held as raw bits (16 B for bf16/fp16, 32 B for fp32)
template <int DT> struct Chunk8 {
    uint4 a;
    uint4 b;  // only used for fp32
};

template <int DT> __device__ __forceinline__ void load_chunk(Chunk8<DT>& c, const void* base, int64_t elem) {
    const char* p = (const char*)base + elem * ElemSize<DT>::value;
    c.a = ldg_stream(p);
    if (DT == DT_F32) c.b = ldg_stream(p + 16);
}
template <int DT> __device__ __forceinline__ void zero_chunk(Chunk8<DT>& c) {
    c.a = make_uint4(0, 0, 0, 0);
    c.b = make_uint4(0, 0, 0, 0);
}
template <int DT> __device__ __forceinline__ void chunk_to_float(const Chunk8<DT>& c, float x[8]) {
    if (DT == DT_BF16) {
        const uint32_t w[4] = {c.a.x, c.a.y, c.a.z, c.a.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            x[2 * i] = __uint_as_float(w[i] << 16);
            x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    } else if (DT == DT_F16) {
        const uint32_t w[4] = {c.a.x, c.a.y, c.a.z, c.a.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            x[2 * i] = f.x;
            x[2 * i + 1] = f.y;
        }
    } else {
        x[0] = __uint_as_float(c.a.x); x[1] = __uint_as_float(c.a.y); x[2] = __uint_as_float(c.a.z); x[3] = __uint_as_float(c.a.w);
        x[4] = __uint_as_float(c.b.x); x[5] = __uint_as_float(c.b.y); x[6] = __uint_as_float(c.b.z); x[7] = __uint_as_float(c.b.w);
    }
}
// 8 floats (already exact in T) -> T bits, stored
template <int DT> __device__ __forceinline__ void store_chunk_T(void* base, int64_t elem, const float y[8]) {
    char* p = (char*)base + elem * ElemSize<DT>::value;
    if (DT == DT_BF16) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        stg_stream(p, make_uint4(w[0], w[1], w[2], w[3]));
    } else if (DT == DT_F16) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            __half2 h = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        stg_stream(p, make_uint4(w[0], w[1], w[2], w[3]));
    } else {
        stg_stream(p, make_uint4(__float_as_uint(y[0]), __float_as_uint(y[1]), __float_as_uint(y[2]), __float_as_uint(y[3])));
        stg_stream(p + 16, make_uint4(__float_as_uint(y[4]), __float_as_uint(y[5]), __float_as_uint(y[6]), __float_as_uint(y[7])));
    }
}

template <int DT> __device__ __forceinline__ float load_T(const void* base, int64_t i) {
    if (DT == DT_BF16) return __uint_as_float(((uint32_t)((const uint16_t*)base)[i]) << 16);
    if (DT == DT_F16) return __half2float(((const __half*)base)[i]);
    return ((const float*)base)[i];
}
template <int DT> __device__ __forceinline__ void store_T(void* base, int64_t i, float v) {
    if (DT == DT_BF16) ((__nv_bfloat16*)base)[i] = __float2bfloat16_rn(v);
    else if (DT == DT_F16) ((__half*)base)[i] = __float2half_rn(v);
    else ((float*)base)[i] = v;
}

// ---- sub-warp butterfly reductions over `lanes` consecutive lanes (power of two, aligned)
__device__ __forceinline__ float subwarp_max(float v, int lanes) {
    for (int o = lanes >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float subwarp_min(float v, int lanes) {
    for (int o = lanes >> 1; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// float atomic max / min through the monotone int mapping (works for any sign, no NaN)
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// dtype dispatch helper for host launchers
#define B200Q_DISPATCH_DT(dt, ...)                                        \
    switch (dt) {                                                         \
    case DT_BF16: { constexpr int DT = DT_BF16; __VA_ARGS__; break; }     \
    case DT_F16: { constexpr int DT = DT_F16; __VA_ARGS__; break; }       \
    case DT_F32: { constexpr int DT = DT_F32; __VA_ARGS__; break; }       \
    default: set_error("unsupported dtype %d", (int)(dt)); return B200Q_EINVAL; \
    }

}  // namespace b200q
