// quant_nvfp4_persistent.cu -- NVFP4 fused compress with the global scale computed in the same launch (bf16):
//     |max| of a sibling span (gate/up of one expert share min(global_scale)) -> global scale -> e4m3 group scales -> e2m1 codes
// (LLMC update_weight_global_scale + update_fused_layer_weight_global_scales + update_weight_zp_scale, then
// CT:compressors/nvfp4/base.py:40-72).  The whole-span |max| must be known before the first code is emitted, i.e. two passes over
// the weight.  Here both passes run in ONE persistent, warp-specialised CTA per SM; HBM sees every weight once, the second read
// is served by the L2 because the two passes are only a few steps of the whole machine (tens of MB at most) apart.
//
//   tile t = blockIdx.x + k * gridDim.x (k = 0, 1, ...: "step" k of this CTA); a span's tiles are consecutive, so the whole
//   machine works on the same few spans at any time.  Per step a CTA runs two jobs through one shared-memory ring of
//   stages (1-D TMA bulk loads):  A(k + D): |max| of tile k + D (the HBM read),  B(k): compress tile k (the L2 re-read).
//   warp 0        producer: bulk-loads the jobs in program order into stage (job % S) once the compute warps released it
//   warp 1        publisher: stores the CTA's |max| of an A job into the tile's global word (value | published flag)
//   warps 2-4     pollers: wait until every tile word of the span of step k is published, take their max, derive the global
//                 scale, build the (e4m3 code -> reciprocal bracket) table of that scale in a ring slot, signal the compute warps
//   other warps   compute: the A / B jobs from shared memory, GPT groups of 16 per thread and job, processed in lock-step phases
//                 (statistics, scale code, conversion, rare exact repair last) so independent groups interleave
// Every CTA runs A(k + D) before B(k), D covers the steps a span stretches over, and A never waits on another CTA, so with
// all CTAs co-resident (cooperative launch) the span of step k is always completed by CTAs that are not themselves blocked: no
// deadlock, no timeouts.  D also gives the publish -> poll -> table chain (~1 us) several steps of slack.
// Arithmetic: the bit-exact chain of qmath.cuh through the bracketed reciprocal of fastmath.cuh (see fp4.cuh).
#include <cstdlib>
#include "async.cuh"
#include "common.cuh"
#include "fastmath.cuh"
#include "fp4.cuh"
#include "kernels.cuh"

namespace b200q {
namespace {
using namespace fast;
using namespace fp4;
using namespace async;

constexpr int RS_Q = 8;                                  // per-step sync slots (ring); must exceed the distance D
constexpr int RS_NPOLL = 3;
constexpr int RS_CTRL = 2 + RS_NPOLL;                    // producer, publisher, pollers
constexpr int RS_MIN_TILE_GROUPS = 512;                  // every configuration's tile holds at least this many groups

template <int NW, int GPT, int S, int CPS = 1> struct RsCfg {
    static constexpr int kCtasPerSm = CPS;
    static constexpr int kTileGroups = NW * 32 * GPT;
    static constexpr int kTileBytes = kTileGroups * 32;
    static constexpr int kThreads = (RS_CTRL + NW) * 32;
    static constexpr int kSmem = S * kTileBytes + RS_Q * 128 * (int)sizeof(Fp4Entry) + (2 * S + 2 * RS_Q) * 8 + RS_Q * 8;
    static_assert(kTileGroups >= RS_MIN_TILE_GROUPS, "workspace sizing assumes tiles of >= 512 groups");
    static_assert(CPS * (kSmem + 1024) <= 233472, "shared memory budget");
};

struct Fp4PersistentParams {
    int64_t groups_per_mat, total_tiles;
    int32_t tiles_per_mat, span_tiles, lookahead;  // lookahead = D
    uint32_t* sync;  // one word per tile: |max| bits | 1 once published; zeroed by the launcher
    float* gs_out;   // [batch]
};

__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// try_wait with a suspend-time hint: the waiting warp sleeps in hardware instead of spinning through the issue slots
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(2000u) : "memory");
}
__device__ __forceinline__ uint32_t absmax2_16(const uint4 a, const uint4 b) {  // packed bf16x2 |max| of 16 elements
    return hmaxabs2(hmaxabs2(hmaxabs2(a.x, a.y), hmaxabs2(a.z, a.w)), hmaxabs2(hmaxabs2(b.x, b.y), hmaxabs2(b.z, b.w)));
}

template <int NW, int GPT, int S, int CPS>
__global__ void __launch_bounds__(RsCfg<NW, GPT, S, CPS>::kThreads, CPS) nvfp4_persistent_kernel(const GroupParams p, const Fp4PersistentParams f) {
    using C = RsCfg<NW, GPT, S, CPS>;
    extern __shared__ __align__(128) uint8_t smem[];
    Fp4Entry* tables = reinterpret_cast<Fp4Entry*>(smem + S * C::kTileBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tables + RS_Q * 128);
    uint32_t* amax_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * RS_Q);
    float* gs_ring = reinterpret_cast<float*>(amax_slot + RS_Q);
    const uint32_t stage0 = smem_u32(smem), bar0 = smem_u32(bars);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (S + s); };
    auto bar_amax = [&](int q) { return bar0 + 8u * (2 * S + q); };
    auto bar_ready = [&](int q) { return bar0 + 8u * (2 * S + RS_Q + q); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t G = gridDim.x, c = blockIdx.x;
    const int K = (int)((f.total_tiles - c + G - 1) / G);   // steps of this CTA
    const int D = f.lookahead;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), NW); }
        for (int q = 0; q < RS_Q; q++) { mbar_init(bar_amax(q), NW); mbar_init(bar_ready(q), 1); amax_slot[q] = 0; }
        fence_async_smem();
    }
    __syncthreads();

    if (warp == 0) {  // ---------------------------------------------------------------- producer
        if (lane == 0) {
            int job = 0;
            auto load_tile = [&](int k) {
                const int s = job % S;
                if (job >= S) mbar_wait_sleepy(bar_empty(s), (uint32_t)(job / S - 1) & 1u);
                const int64_t t = c + (int64_t)k * G, m = t / f.tiles_per_mat, g0 = (t - m * f.tiles_per_mat) * C::kTileGroups;
                const uint32_t bytes = (uint32_t)min((int64_t)C::kTileGroups, f.groups_per_mat - g0) * 32u;
                mbar_arrive_expect_tx(bar_full(s), bytes);
                tma_load_1d(stage0 + s * C::kTileBytes, (const char*)p.w + (m * f.groups_per_mat + g0) * 32, bytes, bar_full(s));
                job++;
            };
            for (int it = -D; it < K; it++) {
                if (it + D < K) load_tile(it + D);
                if (it >= 0) load_tile(it);
            }
        }
    } else if (warp == 1) {  // ---------------------------------------------------------- publisher
        if (lane == 0) {
            for (int k = 0; k < K; k++) {
                const int q = k % RS_Q;
                mbar_wait_sleepy(bar_amax(q), (uint32_t)(k / RS_Q) & 1u);
                const uint32_t bits = *(volatile uint32_t*)&amax_slot[q];
                *(volatile uint32_t*)&amax_slot[q] = 0;
                // one word per tile carries both the value and its "published" flag (bit 0; bf16 bits << 16 leave it free):
                // no fence, no atomic, nothing to order
                st_relaxed(f.sync + (c + (int64_t)k * G), bits | 1u);
            }
        }
    } else if (warp < RS_CTRL) {  // ----------------------------------------------------- pollers
        for (int k = warp - 2; k < K; k += RS_NPOLL) {
            const int q = k % RS_Q;
            const int64_t t = c + (int64_t)k * G, m = t / f.tiles_per_mat;
            const uint32_t* st = f.sync + (t / f.span_tiles) * f.span_tiles;  // the span's tile words
            // tiles are published roughly in order: spin (one lane, with back-off) on the span's last word, then sweep them all
            if (lane == 0) {
                while (ld_relaxed(st + f.span_tiles - 1) == 0) __nanosleep(100);
            }
            __syncwarp();
            uint32_t bits = 0;
            for (int i = lane; i < f.span_tiles; i += 32) {
                uint32_t v = ld_relaxed(st + i);
                while (v == 0) { __nanosleep(50); v = ld_relaxed(st + i); }
                bits = max(bits, v & ~1u);
            }
            bits = __reduce_max_sync(0xffffffffu, bits);
            const float gs = gparam<DT_BF16>(__uint_as_float(bits));
            Fp4Entry* table = tables + q * 128;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int idx = i * 32 + lane;
                // code 0 (scale rounds to zero) is replaced by 0.125 = code 0x20 (helpers.py:101-126); 0x7f is NaN, never produced
                const uint32_t code = idx == 0 ? 0x20u : (uint32_t)idx;
                const float s_eff = fdiv(e4m3_decode((uint8_t)code), gs);
                const float r = rcp_approx(s_eff);
                Fp4Entry e;
                e.r_lo = __fmul_rn(r, 0.99999952316284179688f);
                e.r_hi = __fmul_rn(r, 1.00000047683715820312f);
                e.s_eff = s_eff;
                e.unsafe = fp4_scale_is_safe(s_eff) ? 0.0f : 1.0f;
                table[idx] = e;
            }
            if (lane == 0) {
                gs_ring[q] = gs;
                if (t == m * f.tiles_per_mat) f.gs_out[m] = gs;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready(q));
        }
    } else {  // ------------------------------------------------------------------------- compute
        const int tid = threadIdx.x - RS_CTRL * 32;
        const uint32_t first = (lane >> 2) & 1;   // half read first: keeps each quarter-warp's LDS.128 on 32 distinct banks
        int job = 0;
        for (int it = -D; it < K; it++) {
            const int ka = it + D;
            if (ka < K) {  // A(ka): |max| of the tile
                const int s = job % S;
                mbar_wait(bar_full(s), (uint32_t)(job / S) & 1u);
                job++;
                const int64_t t = c + (int64_t)ka * G, m = t / f.tiles_per_mat, g0 = (t - m * f.tiles_per_mat) * C::kTileGroups;
                const int n = (int)min((int64_t)C::kTileGroups, f.groups_per_mat - g0);
                uint32_t mm = 0;
#pragma unroll
                for (int u = 0; u < GPT; u++) {
                    const int g = u * (NW * 32) + tid;
                    const uint32_t addr = stage0 + s * C::kTileBytes + g * 32;
                    const uint4 v0 = lds128(addr + first * 16), v1 = lds128(addr + (first ^ 1) * 16);
                    if (g < n) mm = hmaxabs2(mm, absmax2_16(v0, v1));  // a partial tile leaves stale bytes behind its end
                }
                mm = hmaxabs2(mm, prmt(mm, mm, 0x1032));
                const uint32_t bits = __reduce_max_sync(0xffffffffu, (mm << 16) & 0x7fff0000u);  // non-negative floats order like uints
                if (lane == 0) {
                    mbar_arrive(bar_empty(s));  // the tile is only read here; its B job loads it again (from the L2)
                    atomicMax(&amax_slot[ka % RS_Q], bits);
                    mbar_arrive(bar_amax(ka % RS_Q));
                }
            }
            if (it >= 0) {  // B(it): compress the tile
                const int s = job % S, q = it % RS_Q;
                mbar_wait(bar_full(s), (uint32_t)(job / S) & 1u);
                job++;
                const int64_t t = c + (int64_t)it * G, m = t / f.tiles_per_mat, g0 = (t - m * f.tiles_per_mat) * C::kTileGroups;
                const int n = (int)min((int64_t)C::kTileGroups, f.groups_per_mat - g0);
                uint8_t* sbase = (uint8_t*)p.scale + m * f.groups_per_mat + g0;
                uint2* obase = reinterpret_cast<uint2*>((uint8_t*)p.out + (m * f.groups_per_mat + g0) * 8);
                // phase 1 (does not need the global scale): loads, group |max|, T(|max| / 6)
                uint4 va[GPT], vb[GPT];  // va: the half read first (elements 8*first .. 8*first+7)
                float loc[GPT];
                uint32_t bad = 0;
#pragma unroll
                for (int u = 0; u < GPT; u++) {
                    const uint32_t addr = stage0 + s * C::kTileBytes + (u * (NW * 32) + tid) * 32;
                    va[u] = lds128(addr + first * 16);
                    vb[u] = lds128(addr + (first ^ 1) * 16);
                }
#pragma unroll
                for (int u = 0; u < GPT; u++) {
                    uint32_t mm = absmax2_16(va[u], vb[u]);
                    mm = hmaxabs2(mm, prmt(mm, mm, 0x1032));
                    const uint32_t abits = (mm << 16) & 0x7fff0000u;
                    const float a = __uint_as_float(abits);
                    const float lo = __fmul_rn(a, (1.0f / 6.0f) * 0.99999952316284179688f), hi = __fmul_rn(a, (1.0f / 6.0f) * 1.00000047683715820312f);
                    const uint32_t w = cvt_bf16x2(hi, lo);
                    const bool in_range = (abits - 0x0d800000u) <= (0x71800000u - 0x0d800000u) || abits == 0;
                    loc[u] = __uint_as_float(w << 16);
                    if (!(in_range && (w >> 16) == (w & 0xffffu))) bad |= 1u << u;
                }
                if (bad) {  // rare: the bracketed constant reciprocal is ambiguous -> IEEE division
#pragma unroll
                    for (int u = 0; u < GPT; u++)
                        if (bad & (1u << u)) {
                            uint32_t mm = absmax2_16(va[u], vb[u]);
                            mm = hmaxabs2(mm, prmt(mm, mm, 0x1032));
                            loc[u] = fp4_loc_scale((mm << 16) & 0x7fff0000u);
                        }
                }
                mbar_wait(bar_ready(q), (uint32_t)(it / RS_Q) & 1u);
                const float gs = gs_ring[q];
                const Fp4Entry* table = tables + q * 128;
                // phase 2: scale code, table fetch, conversion at both bracket ends
                uint32_t code[GPT], pa[GPT], pb[GPT], fix = 0;
                float s_eff[GPT];
#pragma unroll
                for (int u = 0; u < GPT; u++) {
                    const float sf = fminf(__fmul_rn(gs, loc[u]), 448.0f);  // gs * loc, clamp (non-negative)
                    code[u] = cvt_e4m3x2(0.0f, sf) & 0xffu;
                    const Fp4Entry e = table[code[u]];
                    s_eff[u] = e.s_eff;
                    const f32x2 rl = pack2(e.r_lo, e.r_lo), rh = pack2(e.r_hi, e.r_hi);
                    // x * r + 0.0: exact -0.0 inputs carry no sign nibble (torch.sign(-0.) == 0); satfinite == clamp to +-6; the sign
                    // bit of a value that rounds to zero comes from the pre-round sign, like the reference's sign(x) * |q|
                    uint32_t diff;
                    {
                        const f32x2 x0 = bf16x2_to_f32x2(va[u].x), x1 = bf16x2_to_f32x2(va[u].y), x2 = bf16x2_to_f32x2(va[u].z), x3 = bf16x2_to_f32x2(va[u].w);
                        pa[u] = cvt_e2m1x8(mul2_plus0(x0, rl), mul2_plus0(x1, rl), mul2_plus0(x2, rl), mul2_plus0(x3, rl));
                        diff = pa[u] ^ cvt_e2m1x8(mul2_plus0(x0, rh), mul2_plus0(x1, rh), mul2_plus0(x2, rh), mul2_plus0(x3, rh));
                    }
                    {
                        const f32x2 x0 = bf16x2_to_f32x2(vb[u].x), x1 = bf16x2_to_f32x2(vb[u].y), x2 = bf16x2_to_f32x2(vb[u].z), x3 = bf16x2_to_f32x2(vb[u].w);
                        pb[u] = cvt_e2m1x8(mul2_plus0(x0, rl), mul2_plus0(x1, rl), mul2_plus0(x2, rl), mul2_plus0(x3, rl));
                        diff |= pb[u] ^ cvt_e2m1x8(mul2_plus0(x0, rh), mul2_plus0(x1, rh), mul2_plus0(x2, rh), mul2_plus0(x3, rh));
                    }
                    if (diff != 0 || e.unsafe != 0.0f) fix |= 1u << u;
                }
                if (fix) {  // rare: exact per-element repair
#pragma unroll
                    for (int u = 0; u < GPT; u++)
                        if (fix & (1u << u)) {
                            pa[u] = fix_group_fp4(va[u], s_eff[u], pa[u]);
                            pb[u] = fix_group_fp4(vb[u], s_eff[u], pb[u]);
                        }
                }
#pragma unroll
                for (int u = 0; u < GPT; u++) {
                    const int g = u * (NW * 32) + tid;
                    if (g < n) {
                        sbase[g] = (uint8_t)(code[u] ? code[u] : 0x20u);
                        stg_stream(obase + g, first ? make_uint2(pb[u], pa[u]) : make_uint2(pa[u], pb[u]));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty(s));
            }
        }
    }
}

int tune_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

template <int NW, int GPT, int S, int CPS = 1>
int launch_cfg(const GroupParams& p, int64_t batch, int span, float* gs_out, uint32_t* sync, cudaStream_t st, int sms) {
    using C = RsCfg<NW, GPT, S, CPS>;
    const int64_t groups_per_mat = p.rows * (p.cols >> 4);
    const int64_t tiles = (groups_per_mat + C::kTileGroups - 1) / C::kTileGroups;
    const int64_t total = tiles * batch, span_tiles = tiles * span;
    if (tiles >= (1ll << 30) || span_tiles >= (1ll << 30)) return B200Q_ENOSYS;
    const int64_t grid = min((int64_t)sms * CPS, total);
    // D: steps a span can stretch over + slack for the publish -> poll -> table chain
    const int64_t D = (span_tiles + grid - 2) / grid + tune_env("B200Q_FP4_SLACK", 4);
    if (D >= RS_Q) return B200Q_ENOSYS;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(nvfp4_persistent_kernel<NW, GPT, S, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem) != cudaSuccess) {
            cudaGetLastError();
            return B200Q_ENOSYS;
        }
        attr_set = true;
    }
    Fp4PersistentParams f;
    f.groups_per_mat = groups_per_mat;
    f.total_tiles = total;
    f.tiles_per_mat = (int)tiles;
    f.span_tiles = (int)span_tiles;
    f.lookahead = (int)D;
    f.sync = sync;
    f.gs_out = gs_out;
    cudaMemsetAsync(sync, 0, sizeof(uint32_t) * total, st);
    GroupParams pp = p;
    void* args[] = {(void*)&pp, (void*)&f};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)nvfp4_persistent_kernel<NW, GPT, S, CPS>, dim3((unsigned)grid), dim3(C::kThreads), args,
                                                      C::kSmem, st);
    if (e != cudaSuccess) {
        set_error("nvfp4 persistent launch failed: %s", cudaGetErrorString(e));
        return B200Q_ECUDA;
    }
    return B200Q_OK;
}

}  // namespace

// bytes of sync words (one per tile) the persistent kernel needs, for any tile configuration
int64_t nvfp4_resident_workspace(int64_t batch, int64_t rows, int64_t cols) {
    return 4 * batch * ((rows * (cols >> 4) + RS_MIN_TILE_GROUPS - 1) / RS_MIN_TILE_GROUPS);
}

// B200Q_ENOSYS when a span stretches over more steps than the sync ring holds (the caller falls back to the item kernel)
int launch_nvfp4_resident(const GroupParams& p, int64_t batch, int span, float* gs_out, uint32_t* sync, cudaStream_t st) {
    if (p.cols % 16 != 0 || (((uintptr_t)p.w) & 15) != 0 || (((uintptr_t)p.out) & 7) != 0 || batch * p.rows * p.cols == 0) return B200Q_ENOSYS;
    if (span < 1 || batch % span != 0) return B200Q_ENOSYS;
    int dev = 0, sms = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || sms < 1) return B200Q_ENOSYS;
    // measured on B200 (gpurun_out/sweep_nvfp4_e.log): 24 compute warps x 2 groups per thread (48 KB tiles, 4 stages) is the best
    // shape; one group per thread loses ~20 % (no independent chains to interleave)
    switch (tune_env("B200Q_FP4_CFG", 1)) {
    case 0: return launch_cfg<16, 2, 6>(p, batch, span, gs_out, sync, st, sms);
    case 3: return launch_cfg<20, 2, 4>(p, batch, span, gs_out, sync, st, sms);
    case 4: return launch_cfg<12, 2, 3, 2>(p, batch, span, gs_out, sync, st, sms);  // two desynchronised CTAs per SM
    case 5: return launch_cfg<12, 2, 4, 2>(p, batch, span, gs_out, sync, st, sms);
    default: return launch_cfg<24, 2, 4>(p, batch, span, gs_out, sync, st, sms);
    }
}

}  // namespace b200q
