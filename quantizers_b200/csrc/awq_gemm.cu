// awq_gemm.cu -- AWQ loss evaluation for a single-Linear parent on the 5th-gen tensor cores.
//
//   loss[r] += sum_{t,n} ( bf16(X W_ref^T)[t,n] - bf16(X W_q[r]^T)[t,n] )^2        r = 0 .. R-1
//
// (LLMC AWQModifier._run_samples + _compute_loss for the up_proj->down_proj / v->o / per-expert w3->w2 mappings,
// SURVEY.md §8a W2/W3).  The reference materialises every output with cuBLAS, re-reads both for the loss and
// syncs per sample; here the outputs never leave the SM:
//
//   warp 0      TMA producer  : X tile [128 x 64] + W tile [256 x 64] per k-block, SWIZZLE_128B, 4-stage mbarrier ring
//   warp 1      MMA issuer    : tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32), M=128 N=256 K=16, accumulators
//                               in TMEM: columns [0,256) = reference output tile, [256,512) = current ratio's tile
//   warps 2..5  epilogue      : tcgen05.ld the reference tile once per (m,n) tile and keep it as packed bf16 in
//                               registers; per ratio tcgen05.ld the quantised tile, round to bf16, subtract in bf16
//                               (one rounding, like the reference's bf16 tensor subtraction), square-accumulate in
//                               fp32, warp-reduce, one fp64 atomic per warp per (tile, ratio)
//
// Persistent CTAs (one per SM) walk (m-tile, n-tile) items; per item the K loop runs 1 + R times (reference first).
// Tiles out of range are zero-filled by TMA, which contributes exactly 0 to every loss.
//
// Either operand may vary with the ratio: pass 0 multiplies (A_ref, B_ref), pass r >= 1 multiplies
// (A_q[r-1] or A_ref, B_q[r-1] or B_ref).  B varies for a single-Linear parent (fake-quantised weight variants);
// A varies for the last Linear of a multi-layer parent (MLP: A_q[r] = silu(x Wg'_r^T) * (x Wu'_r^T), B = W_down;
// attention: A_q[r] = attention output under the r-th q/k/v variants, B = W_o).
//
// awq_gemm_project_kernel is the companion for the first Linear(s) of such parents: out[v] = bf16(X W[v]^T), or with
// the SwiGLU epilogue out[v] = silu(bf16(X Wg[v]^T)) * bf16(X Wu[v]^T) (gate tile in TMEM columns [0,128), up tile in
// [128,256) of the same accumulator), double-buffered TMEM so the epilogue of one tile overlaps the MMAs of the next.
#include <cuda.h>
#include <cudaTypedefs.h>
#include "../../include/b200q.h"
#include "common.cuh"

namespace b200q {
namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2, kBBytes = BN * BK * 2, kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float hi, float lo) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// K-major operand tile, 128-byte rows, SWIZZLE_128B, 8-row groups 1024 B apart (sm_100 shared-memory descriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D=f32 (bit 4), A=bf16 (bit 7), B=bf16 (bit 10), both K-major, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct GemmParams {
    int32_t n_m_tiles, n_n_tiles, n_k_blocks, n_ratios, n_fastest, a_varies, b_varies;
    double* acc;  // [n_ratios]
};

__global__ void __launch_bounds__(kThreads, 1)
awq_gemm_loss_kernel(const __grid_constant__ CUtensorMap map_aref, const __grid_constant__ CUtensorMap map_aq,
                     const __grid_constant__ CUtensorMap map_bref, const __grid_constant__ CUtensorMap map_bq, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t tiles = smem_u32(smem);
    const uint32_t bars = tiles + kStages * kStageBytes;
    // barrier map: full[s] = bars + 8s ; empty[s] = bars + 8(kStages + s) ; tfull[i] = +8(2kStages + i) ; tempty[i] = +8(2kStages + 2 + i)
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_tfull = bars + 16 * kStages, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_slot = (uint32_t*)(smem + kStages * kStageBytes + 16 * kStages + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int i = 0; i < 2; i++) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.n_m_tiles * p.n_n_tiles;
    const int passes = p.n_ratios + 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int mt = p.n_fastest ? item / p.n_n_tiles : item % p.n_m_tiles;
                const int nt = p.n_fastest ? item % p.n_n_tiles : item / p.n_m_tiles;
                for (int pass = 0; pass < passes; pass++) {
                    for (int kb = 0; kb < p.n_k_blocks; kb++, it++) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                        mbar_arrive_expect_tx(bar_full + 8 * s, kStageBytes);
                        const uint32_t a_dst = tiles + s * kStageBytes, b_dst = a_dst + kABytes;
                        if (pass == 0 || !p.a_varies) tma_load_2d(a_dst, &map_aref, kb * BK, mt * BM, bar_full + 8 * s);
                        else tma_load_3d(a_dst, &map_aq, kb * BK, mt * BM, pass - 1, bar_full + 8 * s);
                        if (pass == 0 || !p.b_varies) tma_load_2d(b_dst, &map_bref, kb * BK, nt * BN, bar_full + 8 * s);
                        else tma_load_3d(b_dst, &map_bq, kb * BK, nt * BN, pass - 1, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            // Accumulator ping-pong: pass 0 (reference) lands in slot 0 and is parked in the epilogue's registers right away, so
            // TMEM columns [0, 256) are idle for the remaining R passes -- the ratio passes alternate slot 1, 0, 1, ... and the
            // MMAs of ratio r + 1 run while the epilogue warps still drain ratio r (round 1 had every ratio on slot 1: MMA and
            // epilogue serialised on one accumulator).
            uint32_t it = 0, use[2] = {0, 0};
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                for (int pass = 0; pass < passes; pass++) {
                    const uint32_t slot = (uint32_t)pass & 1u;
                    uint32_t& uses = use[slot];
                    mbar_wait(bar_tempty + 8 * slot, (uses & 1u) ^ 1u);  // epilogue has drained this accumulator
                    uses++;
                    tc_fence_after();
                    const uint32_t d = tmem_base + slot * BN;
                    for (int kb = 0; kb < p.n_k_blocks; kb++, it++) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                        mbar_wait(bar_full + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t a_addr = tiles + s * kStageBytes, b_addr = a_addr + kABytes;
                        const uint64_t ad = make_desc(a_addr), bd = make_desc(b_addr);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) tc_mma(d, ad + 2 * k, bd + 2 * k, kIdesc, (kb | k) ? 1u : 0u);
                        tc_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
                    }
                    tc_commit(bar_tfull + 8 * slot);   // accumulator complete
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5 -> TMEM lane quadrant warp & 3)
        const uint32_t quad = (uint32_t)warp & 3u;
        const uint32_t t_lane = tmem_base + ((quad * 32u) << 16);
        uint32_t use[2] = {0, 0};
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t ref[BN / 2];  // this thread's output row of the reference tile, packed bf16x2
            mbar_wait(bar_tfull + 0, use[0] & 1u);
            use[0]++;
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                tc_ld32(t_lane + c * 32, v);
#pragma unroll
                for (int j = 0; j < 16; j++) ref[c * 16 + j] = cvt_bf16x2(__uint_as_float(v[2 * j + 1]), __uint_as_float(v[2 * j]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 0);
            for (int r = 0; r < p.n_ratios; r++) {
                const uint32_t slot = (uint32_t)(r + 1) & 1u;  // pass r + 1 of the MMA issuer's slot sequence
                mbar_wait(bar_tfull + 8 * slot, use[slot] & 1u);
                use[slot]++;
                tc_fence_after();
                float part = 0.0f;
#pragma unroll
                for (int c = 0; c < BN / 32; c++) {
                    uint32_t v[32];
                    tc_ld32(t_lane + slot * BN + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const uint32_t y = cvt_bf16x2(__uint_as_float(v[2 * j + 1]), __uint_as_float(v[2 * j]));
                        const uint32_t d = hsub2(ref[c * 16 + j], y);  // bf16 difference, one rounding
                        const float dl = __uint_as_float(d << 16), dh = __uint_as_float(d & 0xffff0000u);
                        part = fmaf(dl, dl, part);
                        part = fmaf(dh, dh, part);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (lane == 0) atomicAdd(&p.acc[r], (double)part);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// =============================================================================================================
// Projection of all weight variants: out[v, t, n] = bf16(sum_k X[t,k] W[v,n,k])            (SWIGLU = false)
//                                    out[v, t, n] = silu(bf16(X Wg[v]^T)) * bf16(X Wu[v]^T)  (SWIGLU = true; W[v] = [Wg; Wu])
// =============================================================================================================
struct ProjParams {
    int32_t n_m_tiles, n_n_tiles, n_k_blocks, n_variants, group_m;
    int32_t tokens, n_out, up_row_offset;
    uint16_t* out;  // bf16 [n_variants, tokens, n_out]
    // grouped mode (MoE experts): non-null -> one weight variant PER M TILE (tile_variant[mt] = expert owning the tile's rows, < 0 =
    // padding tile, skipped) and a single [tokens, n_out] output; the rows of an expert are padded to whole tiles by the caller
    const int32_t* tile_variant;
    float alpha, beta;  // EPI_F32_ACC: out(fp32) = alpha * out + beta * acc
};

// grouped rasterisation: consecutive items cover group_m m-tiles x all n-tiles column by column, so one wave of 148
// CTAs touches ~group_m X tiles and ~148/group_m W tiles (L2-sized working set) instead of a full row or column
__device__ __forceinline__ void decode_item(const ProjParams& p, int item, int& v, int& mt, int& nt) {
    const int per_variant = p.n_m_tiles * p.n_n_tiles;
    v = item / per_variant;
    const int rem = item - v * per_variant;
    const int gsz = p.group_m * p.n_n_tiles;
    const int grp = rem / gsz;
    const int first_m = grp * p.group_m;
    const int rows = min(p.group_m, p.n_m_tiles - first_m);
    const int within = rem - grp * gsz;
    mt = first_m + within % rows;
    nt = within / rows;
}

// torch.nn.functional.silu on a bf16 tensor: opmath float, x / (1 + exp(-x)), rounded to bf16
__device__ __forceinline__ float silu_f32(float x) { return __fdiv_rn(x, __fadd_rn(1.0f, expf(-x))); }
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

enum : int { EPI_BF16 = 0, EPI_SWIGLU = 1, EPI_F32_ACC = 2 };

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
awq_gemm_project_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const ProjParams p) {
    constexpr bool SWIGLU = EPI == EPI_SWIGLU;
    constexpr int TN = SWIGLU ? BN / 2 : BN;  // output columns per tile
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t tiles = smem_u32(smem);
    const uint32_t bars = tiles + kStages * kStageBytes;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_tfull = bars + 16 * kStages, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_slot = (uint32_t*)(smem + kStages * kStageBytes + 16 * kStages + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int i = 0; i < 2; i++) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool grouped = p.tile_variant != nullptr;
    const int n_items = (grouped ? 1 : p.n_variants) * p.n_m_tiles * p.n_n_tiles;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int v, mt, nt;
                decode_item(p, item, v, mt, nt);
                if (grouped) { v = p.tile_variant[mt]; if (v < 0) continue; }
                for (int kb = 0; kb < p.n_k_blocks; kb++, it++) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    mbar_arrive_expect_tx(bar_full + 8 * s, kStageBytes);
                    const uint32_t a_dst = tiles + s * kStageBytes, b_dst = a_dst + kABytes;
                    tma_load_2d(a_dst, &map_x, kb * BK, mt * BM, bar_full + 8 * s);
                    if (SWIGLU) {
                        tma_load_3d(b_dst, &map_w, kb * BK, nt * TN, v, bar_full + 8 * s);                               // gate rows
                        tma_load_3d(b_dst + kBBytes / 2, &map_w, kb * BK, p.up_row_offset + nt * TN, v, bar_full + 8 * s);  // up rows
                    } else {
                        tma_load_3d(b_dst, &map_w, kb * BK, nt * TN, v, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, n_done = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                if (grouped) {
                    int v, mt, nt;
                    decode_item(p, item, v, mt, nt);
                    if (p.tile_variant[mt] < 0) continue;
                }
                const uint32_t slot = n_done & 1u, use = n_done >> 1;
                n_done++;
                mbar_wait(bar_tempty + 8 * slot, (use & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + slot * BN;
                for (int kb = 0; kb < p.n_k_blocks; kb++, it++) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t a_addr = tiles + s * kStageBytes, b_addr = a_addr + kABytes;
                    const uint64_t ad = make_desc(a_addr), bd = make_desc(b_addr);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) tc_mma(d, ad + 2 * k, bd + 2 * k, kIdesc, (kb | k) ? 1u : 0u);
                    tc_commit(bar_empty + 8 * s);
                }
                tc_commit(bar_tfull + 8 * slot);
            }
        }
    } else {
        const uint32_t quad = (uint32_t)warp & 3u;
        uint32_t n_done = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int v, mt, nt;
            decode_item(p, item, v, mt, nt);
            if (grouped) { if (p.tile_variant[mt] < 0) continue; v = 0; }
            const uint32_t slot = n_done & 1u, use = n_done >> 1;
            n_done++;
            const uint32_t t_lane = tmem_base + ((quad * 32u) << 16) + slot * BN;
            const int row = mt * BM + (int)quad * 32 + lane;
            const int col0 = nt * TN;
            uint16_t* orow = p.out + ((size_t)v * (size_t)p.tokens + (size_t)row) * (size_t)p.n_out + col0;
            mbar_wait(bar_tfull + 8 * slot, use & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < TN / 32; c++) {
                uint32_t g[32], o[16];
                tc_ld32(t_lane + c * 32, g);
                if (EPI == EPI_F32_ACC) {  // fp32 read-modify-write of the caller's matrix (GPTQ Hessian accumulation)
                    if (row < p.tokens) {
                        float* hrow = reinterpret_cast<float*>(p.out) + (size_t)row * (size_t)p.n_out + col0 + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (col0 + c * 32 + j < p.n_out) hrow[j] = __fmaf_rn(p.alpha, hrow[j], __fmul_rn(p.beta, __uint_as_float(g[j])));
                    }
                    continue;
                }
                if (SWIGLU) {
                    uint32_t u[32];
                    tc_ld32(t_lane + TN + c * 32, u);
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float h[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const float gb = bf16_round(__uint_as_float(g[2 * j + e]));
                            const float ub = bf16_round(__uint_as_float(u[2 * j + e]));
                            h[e] = __fmul_rn(bf16_round(silu_f32(gb)), ub);
                        }
                        o[j] = cvt_bf16x2(h[1], h[0]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) o[j] = cvt_bf16x2(__uint_as_float(g[2 * j + 1]), __uint_as_float(g[2 * j]));
                }
                if (row < p.tokens) {
                    const int cbase = col0 + c * 32;
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        if (cbase + q4 * 8 + 8 <= p.n_out) {
                            *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = make_uint4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
                        } else {
                            for (int e = 0; e < 8; e++)
                                if (cbase + q4 * 8 + e < p.n_out) orow[c * 32 + q4 * 8 + e] = (uint16_t)(o[4 * q4 + (e >> 1)] >> (16 * (e & 1)));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

__global__ void finalize_loss_kernel(const double* acc, float* loss, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) loss[i] += (float)acc[i];
}

PFN_cuTensorMapEncodeTiled get_encode() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled)ptr;
    }
    return fn;
}

int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return B200Q_ECUDA; }
    const uint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return B200Q_ECUDA; }
    return B200Q_OK;
}

}  // namespace
}  // namespace b200q

using namespace b200q;

namespace {
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + 16 * kStages + 64 + 1024;

int map_2d(CUtensorMap* m, const void* base, int64_t rows, int64_t k, int box_rows) {
    const uint64_t d[2] = {(uint64_t)k, (uint64_t)rows}, s[1] = {(uint64_t)k * 2};
    const uint32_t b[2] = {BK, (uint32_t)box_rows};
    return make_map(m, base, 2, d, s, b);
}
int map_3d(CUtensorMap* m, const void* base, int64_t count, int64_t rows, int64_t k, int box_rows) {
    const uint64_t d[3] = {(uint64_t)k, (uint64_t)rows, (uint64_t)count}, s[2] = {(uint64_t)k * 2, (uint64_t)k * 2 * (uint64_t)rows};
    const uint32_t b[3] = {BK, (uint32_t)box_rows, 1};
    return make_map(m, base, 3, d, s, b);
}
}  // namespace

extern "C" {

int64_t b200q_awq_gemm_loss_workspace(int64_t, int64_t, int64_t, int32_t n_ratios) { return (int64_t)sizeof(double) * (n_ratios > 0 ? n_ratios : 1); }

int b200q_awq_gemm_loss_pairs(const void* a_ref, const void* a_q, int64_t tokens, int64_t k, const void* b_ref, const void* b_q,
                              int64_t n, int32_t n_ratios, float* loss, void* workspace, int64_t workspace_bytes, void* stream) {
    B200Q_REQUIRE(a_ref && b_ref && loss && workspace, "b200q_awq_gemm_loss: NULL pointer");
    B200Q_REQUIRE(a_q || b_q, "b200q_awq_gemm_loss: neither operand varies with the ratio");
    B200Q_REQUIRE(n_ratios >= 1 && n_ratios <= 1024, "n_ratios out of range");
    B200Q_REQUIRE(workspace_bytes >= (int64_t)sizeof(double) * n_ratios, "workspace too small");
    B200Q_REQUIRE(k % 8 == 0 && k >= 8, "K must be a multiple of 8 (16-byte rows for the tensor maps), got %lld", (long long)k);
    B200Q_REQUIRE(tokens >= 1 && n >= 1, "empty problem");
    B200Q_REQUIRE((((uintptr_t)a_ref | (uintptr_t)a_q | (uintptr_t)b_ref | (uintptr_t)b_q) & 15) == 0, "operands must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap ma, maq, mb, mbq;
    if (int rc = map_2d(&ma, a_ref, tokens, k, BM)) return rc;
    if (int rc = map_2d(&mb, b_ref, n, k, BN)) return rc;
    if (int rc = a_q ? map_3d(&maq, a_q, n_ratios, tokens, k, BM) : map_2d(&maq, a_ref, tokens, k, BM)) return rc;
    if (int rc = b_q ? map_3d(&mbq, b_q, n_ratios, n, k, BN) : map_2d(&mbq, b_ref, n, k, BN)) return rc;
    GemmParams p;
    p.n_m_tiles = (int)((tokens + BM - 1) / BM);
    p.n_n_tiles = (int)((n + BN - 1) / BN);
    p.n_k_blocks = (int)((k + BK - 1) / BK);
    p.n_ratios = n_ratios;
    p.a_varies = a_q ? 1 : 0;
    p.b_varies = b_q ? 1 : 0;
    p.acc = (double*)workspace;
    // keep the operand that is re-streamed (1 + R) times L2-resident.  A varies: the CTAs of one m-tile run side by side
    // (n fastest) so every A_q[r] tile is fetched from HBM once.  B varies: with a long K the 148 concurrent A tiles do not
    // fit in L2, so neighbours share an A tile (n fastest); otherwise they share the B tile (m fastest)
    p.n_fastest = (p.a_varies || (int64_t)kNumSMs * BM * k * 2 > (64ll << 20)) ? 1 : 0;
    cudaMemsetAsync(workspace, 0, sizeof(double) * n_ratios, st);
    cudaFuncSetAttribute(awq_gemm_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
    const int grid = (int)min((int64_t)kNumSMs, (int64_t)p.n_m_tiles * p.n_n_tiles);
    awq_gemm_loss_kernel<<<grid, kThreads, kGemmSmem, st>>>(ma, maq, mb, mbq, p);
    B200Q_CHECK_LAUNCH();
    finalize_loss_kernel<<<(n_ratios + 127) / 128, 128, 0, st>>>(p.acc, loss, n_ratios);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

int b200q_awq_gemm_loss(const void* x, int64_t tokens, int64_t k, const void* w_ref, const void* w_q, int64_t n, int32_t n_ratios,
                        float* loss, void* workspace, int64_t workspace_bytes, void* stream) {
    B200Q_REQUIRE(w_q, "b200q_awq_gemm_loss: NULL pointer");
    return b200q_awq_gemm_loss_pairs(x, nullptr, tokens, k, w_ref, w_q, n, n_ratios, loss, workspace, workspace_bytes, stream);
}

static int gemm_project_impl(const void* x, int64_t tokens, int64_t k, const void* w, int64_t n_variants, int64_t n_out, int32_t swiglu,
                             const void* tile_variant, void* out, void* stream, int epi_f32 = 0, float alpha = 0.0f, float beta = 1.0f) {
    B200Q_REQUIRE(x && w && out, "b200q_awq_gemm_project: NULL pointer");
    B200Q_REQUIRE(k % 8 == 0 && k >= 8, "K must be a multiple of 8 (16-byte rows for the tensor maps), got %lld", (long long)k);
    B200Q_REQUIRE(n_out % 8 == 0, "n_out must be a multiple of 8 (16-byte output vectors), got %lld", (long long)n_out);
    B200Q_REQUIRE(tokens >= 1 && n_out >= 1 && n_variants >= 1, "empty problem");
    B200Q_REQUIRE(tokens < (1ll << 31) && n_out < (1ll << 30), "problem too large for 32-bit tile indices");
    B200Q_REQUIRE((((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) & 15) == 0, "operands must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int tn = swiglu ? BN / 2 : BN;
    const int64_t w_rows = swiglu ? 2 * n_out : n_out;
    CUtensorMap mx, mw;
    if (int rc = map_2d(&mx, x, tokens, k, BM)) return rc;
    if (int rc = map_3d(&mw, w, n_variants, w_rows, k, tn)) return rc;
    ProjParams p;
    p.n_m_tiles = (int)((tokens + BM - 1) / BM);
    p.n_n_tiles = (int)((n_out + tn - 1) / tn);
    p.n_k_blocks = (int)((k + BK - 1) / BK);
    p.n_variants = (int)n_variants;
    p.group_m = 16;
    p.tokens = (int)tokens;
    p.n_out = (int)n_out;
    p.up_row_offset = (int)n_out;
    p.out = (uint16_t*)out;
    p.tile_variant = (const int32_t*)tile_variant;
    p.alpha = alpha;
    p.beta = beta;
    const int64_t items = (int64_t)(tile_variant ? 1 : p.n_variants) * p.n_m_tiles * p.n_n_tiles;
    B200Q_REQUIRE(items < (1ll << 31), "too many tiles");
    const int grid = (int)min((int64_t)kNumSMs, items);
    if (epi_f32) {
        cudaFuncSetAttribute(awq_gemm_project_kernel<EPI_F32_ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
        awq_gemm_project_kernel<EPI_F32_ACC><<<grid, kThreads, kGemmSmem, st>>>(mx, mw, p);
    } else if (swiglu) {
        cudaFuncSetAttribute(awq_gemm_project_kernel<EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
        awq_gemm_project_kernel<EPI_SWIGLU><<<grid, kThreads, kGemmSmem, st>>>(mx, mw, p);
    } else {
        cudaFuncSetAttribute(awq_gemm_project_kernel<EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
        awq_gemm_project_kernel<EPI_BF16><<<grid, kThreads, kGemmSmem, st>>>(mx, mw, p);
    }
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

int b200q_awq_gemm_project(const void* x, int64_t tokens, int64_t k, const void* w, int64_t n_variants, int64_t n_out, int32_t swiglu,
                           void* out, void* stream) {
    return gemm_project_impl(x, tokens, k, w, n_variants, n_out, swiglu, nullptr, out, stream);
}

int b200q_gptq_hessian_accumulate(const void* xt, int64_t features, int64_t tokens, float alpha, float beta, float* hessian, void* stream) {
    B200Q_REQUIRE(hessian, "b200q_gptq_hessian_accumulate: NULL pointer");
    B200Q_REQUIRE((((uintptr_t)hessian) & 15) == 0, "hessian must be 16-byte aligned");
    // H = alpha * H + beta * (X^T X): both operands are X^T [features, tokens] (contraction over the tokens, which are contiguous)
    return gemm_project_impl(xt, features, tokens, xt, 1, features, 0, nullptr, hessian, stream, 1, alpha, beta);
}

int b200q_awq_gemm_project_grouped(const void* x, int64_t tokens, int64_t k, const void* w, int64_t n_experts, int64_t n_out, int32_t swiglu,
                                   const int32_t* tile_expert, void* out, void* stream) {
    B200Q_REQUIRE(tile_expert, "b200q_awq_gemm_project_grouped: NULL tile table");
    B200Q_REQUIRE(tokens % BM == 0, "grouped projection: the rows of every expert are padded to whole %d-row tiles by the caller", BM);
    return gemm_project_impl(x, tokens, k, w, n_experts, n_out, swiglu, tile_expert, out, stream);
}

}  // extern "C"
