// awq_attn_core.cu -- causal grouped-query attention core of the AWQ attention parent (LLMC AWQModifier._run_samples on
// ``self_attn``; transformers Qwen3Attention: softmax(Q K^T / sqrt(d), causal) V between the q/k/v projections and o_proj), on the
// 5th-gen tensor cores.  Round 1 called torch's scaled_dot_product_attention here (a library kernel on the hot path, 21 x per
// q/k/v mapping); this is the hand-written replacement.
//
// Inputs are the projected rows that awq_gemm_project_kernel wrote and b200q_qk_norm_rope normalised / rotated in place:
//   qkv [T, (H + 2 Hkv) d] bf16, T = samples * seq_len; every sample attends within itself (batch-1 forwards in the reference).
//   vt  [samples * Hkv * d, S_pad] bf16: V transposed per (sample, kv head) by attn_transpose_v_kernel, so that BOTH GEMMs take
//   K-major operands (the same SWIZZLE_128B tiles / descriptors as awq_gemm.cu): S = Q K^T contracts over d, O = P V over keys.
//
// Work item = (sample, q head, 128-query block); key blocks 0 .. qb (causal); persistent CTAs, one per SM:
//   warp 0      TMA producer : Q once; per key block K_j [128 keys x d] and V^T_j [d x 128 keys] into a 2-stage ring
//   warp 1      MMA issuer   : S_{j+1} = Q K_{j+1}^T is issued BEFORE O += P_j V_j, so the softmax of block j overlaps it
//                              (S double-buffered in TMEM columns [0,128) / [128,256), O in [256, 256 + d))
//   warps 2..5  softmax      : one query row per thread (TMEM lane = row): the 128 scores of the row in registers (four x32
//                              tcgen05.ld in flight, one wait), scaled / masked row max, LAZY running maximum (moves only when
//                              the row max grew by > 2^8, so O in TMEM is almost never rescaled: tcgen05.ld / st only then),
//                              exp2 -> row sum, P as bf16 into the SWIZZLE_128B A tile in shared memory; epilogue O / l -> bf16.
// fp32 scores, fp32 softmax statistics, bf16 probabilities, fp32 accumulation -- the arithmetic of a flash-attention forward.
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "../../include/b200q.h"
#include "common.cuh"

namespace b200q {
namespace {

constexpr int BQ = 128, BKEY = 128, SUBK = 64;      // query block, key block, 128-byte (64 x bf16) K-major sub-tile width
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;
constexpr int kSubBytes = 128 * SUBK * 2;           // one [128 rows x 64] sub-tile = 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// (a spin on mbarrier.test_wait instead of try_wait was measured: no difference, 440 vs 427 us)
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
// the same load without the wait: issue several, then tc_wait_ld() once
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t cvt_bf16x2(float hi, float lo) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// K-major operand tile, 128-byte rows, SWIZZLE_128B, 8-row groups 1024 B apart (same descriptor as awq_gemm.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N as given
__device__ __forceinline__ constexpr uint32_t idesc_n(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}

struct AttnParams {
    int32_t n_heads, n_kv, seq_len, q_blocks;   // q_blocks = ceil(seq_len / 128)
    int32_t samples;
    float scale_log2e;                          // softmax scale * log2(e)
    float lazy_raw;                             // 8 / scale_log2e: the lazy-maximum head-room (2^8) in raw score units
    uint16_t* out;                              // bf16 [T, n_heads * D]
};

// ------------------------------------------------------------------------------------------------ V -> V^T per (sample, kv head)
__global__ void attn_transpose_v_kernel(const uint16_t* __restrict__ qkv, int64_t row_elems, int v_col0, int d, int seq_len, int s_pad,
                                        int n_kv, uint16_t* __restrict__ vt) {
    __shared__ uint16_t tile[32][34];
    const int bh = blockIdx.z, b = bh / n_kv, hk = bh - b * n_kv;
    const int k0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int key = k0 + ty + 8 * i;
        tile[ty + 8 * i][tx] = key < seq_len ? qkv[((int64_t)b * seq_len + key) * row_elems + v_col0 + hk * d + d0 + tx] : (uint16_t)0;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int dd = d0 + ty + 8 * i, key = k0 + tx;
        if (key < s_pad) vt[((int64_t)bh * d + dd) * s_pad + key] = tile[tx][ty + 8 * i];
    }
}

// ------------------------------------------------------------------------------------------------ attention core
// Persistent: one CTA per SM walks the (query block, sample, head) items -- longest (latest query block) first -- and all three
// roles run one item ahead of each other: Q is double-buffered in shared memory, O in TMEM (columns [256, 256 + D) / [384, 384 + D)),
// and the MMA warp issues the first S = Q K^T of the NEXT item before the last O += P V of the current one, so TMA latency, the
// softmax and the epilogue of neighbouring items overlap (the first cut, one CTA per item, spent ~6 of its ~10 us per item on
// exposed start-up / drain: 554 us for the AWQ shape vs 217 us for cuDNN's flash kernel).
struct ItemIter {
    int t, end, stride;
    int q_blocks, heads_total, n_heads, n_kv, seq_len;
    __device__ __forceinline__ bool valid() const { return t < end; }
    __device__ __forceinline__ void next() { t += stride; }
    __device__ __forceinline__ void decode(int& b, int& h, int& qb, int& n_blocks) const {
        qb = q_blocks - 1 - t / heads_total;          // items sorted by length: all (sample, head) pairs of the last query block first
        const int bh = t - (t / heads_total) * heads_total;
        b = bh / n_heads;
        h = bh - b * n_heads;
        n_blocks = min(qb + 1, (seq_len + BKEY - 1) / BKEY);
    }
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
attn_core_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_vt, const AttnParams p) {
    constexpr int NSUB = D / SUBK;                          // 64-wide sub-tiles along d
    constexpr int kQBytes = NSUB * kSubBytes;               // Q tile / K tile: [128 x D]
    constexpr int kVSub = D * SUBK * 2;                     // V^T sub-tile: [D rows x 64 keys]
    constexpr int kVBytes = 2 * kVSub;                      // V^T tile: [D x 128 keys]
    constexpr int kPBytes = 2 * kSubBytes;                  // P tile: [128 x 128 keys]
    constexpr int kStageBytes = kQBytes + kVBytes;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t q_s = smem_u32(smem);                    // Q buffer i at q_s + i * kQBytes
    const uint32_t kv_s = q_s + 2 * kQBytes;                // stage s: K at kv_s + s * kStageBytes, V^T right after it
    const uint32_t p_s = kv_s + 2 * kStageBytes;
    const uint32_t bars = p_s + kPBytes;
    // barriers: q_full[2], q_empty[2], kv_full[2], kv_empty[2], s_full[2], s_empty[2], o_empty[2], p_full, p_free
    const uint32_t bar_qf = bars, bar_qe = bars + 16, bar_kvf = bars + 32, bar_kve = bars + 48, bar_sf = bars + 64, bar_se = bars + 80,
                   bar_oe = bars + 96, bar_pf = bars + 112, bar_pfree = bars + 120;
    uint32_t* tmem_slot = (uint32_t*)(smem + 2 * kQBytes + 2 * kStageBytes + kPBytes + 136);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_qf + 8 * s, 1); mbar_init(bar_qe + 8 * s, 1); mbar_init(bar_kvf + 8 * s, 1); mbar_init(bar_kve + 8 * s, 1);
            mbar_init(bar_sf + 8 * s, 1); mbar_init(bar_se + 8 * s, 4); mbar_init(bar_oe + 8 * s, 4);
        }
        mbar_init(bar_pf, 4);
        mbar_init(bar_pfree, 1);
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

    ItemIter iter;
    iter.t = (int)blockIdx.x; iter.stride = (int)gridDim.x;
    iter.heads_total = p.samples * p.n_heads; iter.end = iter.heads_total * p.q_blocks;
    iter.q_blocks = p.q_blocks; iter.n_heads = p.n_heads; iter.n_kv = p.n_kv; iter.seq_len = p.seq_len;
    const int kv_rep = p.n_heads / p.n_kv;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0, g = 0;
            for (; iter.valid(); iter.next(), it++) {
                int b, h, qb, nb;
                iter.decode(b, h, qb, nb);
                const int hk = h / kv_rep, row0 = b * p.seq_len, qbuf = (int)(it & 1u);
                mbar_wait(bar_qe + 8 * qbuf, ((it >> 1) & 1u) ^ 1u);           // the S MMAs of item it - 2 have retired
                mbar_arrive_expect_tx(bar_qf + 8 * qbuf, kQBytes);
#pragma unroll
                for (int t = 0; t < NSUB; t++) tma_load_2d(q_s + qbuf * kQBytes + t * kSubBytes, &map_qk, h * D + t * SUBK, row0 + qb * BQ, bar_qf + 8 * qbuf);
                for (int j = 0; j < nb; j++, g++) {
                    const uint32_t s = g & 1u;
                    mbar_wait(bar_kve + 8 * s, ((g >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_kvf + 8 * s, kStageBytes);
                    const uint32_t k_dst = kv_s + s * kStageBytes, v_dst = k_dst + kQBytes;
#pragma unroll
                    for (int t = 0; t < NSUB; t++)
                        tma_load_2d(k_dst + t * kSubBytes, &map_qk, (p.n_heads + hk) * D + t * SUBK, row0 + j * BKEY, bar_kvf + 8 * s);
#pragma unroll
                    for (int t = 0; t < 2; t++)
                        tma_load_2d(v_dst + t * kVSub, &map_vt, j * BKEY + t * SUBK, (b * p.n_kv + hk) * D, bar_kvf + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: a flat stream of key blocks, S one block ahead of P V
        if (lane == 0) {
            // "ahead" cursor (issues S) and "behind" cursor (issues P V) over the same item sequence
            ItemIter ia = iter, ib = iter;
            uint32_t it_a = 0, g_a = 0, it_b = 0, g_b = 0;
            int ja = 0, nba = 0, jb = 0, nbb = 0;
            bool a_open = false;                                   // cursor a positioned inside an item?
            auto advance_a = [&]() -> bool {                        // issue S of the next block of the stream; false when the stream is exhausted
                if (!a_open) {
                    if (!ia.valid()) return false;
                    int b, h, qb;
                    ia.decode(b, h, qb, nba);
                    ja = 0;
                    a_open = true;
                    mbar_wait(bar_qf + 8 * (it_a & 1u), (it_a >> 1) & 1u);
                }
                const uint32_t s = g_a & 1u, qbuf = it_a & 1u;
                mbar_wait(bar_kvf + 8 * s, (g_a >> 1) & 1u);
                mbar_wait(bar_se + 8 * s, ((g_a >> 1) & 1u) ^ 1u);  // softmax has drained this S slot
                tc_fence_after();
                const uint32_t k_addr = kv_s + s * kStageBytes, q_addr = q_s + qbuf * kQBytes;
#pragma unroll
                for (int k = 0; k < D / 16; k++) {
                    const uint64_t ad = make_desc(q_addr + (k / 4) * kSubBytes) + 2 * (k % 4);
                    const uint64_t bd = make_desc(k_addr + (k / 4) * kSubBytes) + 2 * (k % 4);
                    tc_mma(tmem_base + s * BKEY, ad, bd, idesc_n(BKEY), k ? 1u : 0u);
                }
                tc_commit(bar_sf + 8 * s);
                g_a++;
                if (++ja == nba) {                                  // last S of the item: its Q buffer is free once these MMAs retire
                    tc_commit(bar_qe + 8 * qbuf);
                    a_open = false;
                    ia.next();
                    it_a++;
                }
                return true;
            };
            advance_a();
            for (; ib.valid(); ib.next(), it_b++) {
                int b, h, qb;
                ib.decode(b, h, qb, nbb);
                const uint32_t obuf = it_b & 1u;
                const uint32_t tmem_o = tmem_base + 256 + obuf * 128;
                for (jb = 0; jb < nbb; jb++, g_b++) {
                    advance_a();                                    // S of block g_b + 1 (possibly the next item's first block)
                    const uint32_t s = g_b & 1u;
                    mbar_wait(bar_pf, g_b & 1u);                    // P ready in shared memory, O rescaled
                    if (jb == 0) mbar_wait(bar_oe + 8 * obuf, ((it_b >> 1) & 1u) ^ 1u);   // epilogue of item it_b - 2 has read this O buffer
                    tc_fence_after();
                    const uint32_t v_addr = kv_s + s * kStageBytes + kQBytes;
#pragma unroll
                    for (int k = 0; k < BKEY / 16; k++) {
                        const uint64_t ad = make_desc(p_s + (k / 4) * kSubBytes) + 2 * (k % 4);
                        const uint64_t bd = make_desc(v_addr + (k / 4) * kVSub) + 2 * (k % 4);
                        tc_mma(tmem_o, ad, bd, idesc_n(D), (jb | k) ? 1u : 0u);
                    }
                    tc_commit(bar_kve + 8 * s);      // K / V stage free
                    tc_commit(bar_pfree);            // P buffer free, O updated
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax + epilogue: one query row per thread
        const uint32_t quad = (uint32_t)warp & 3u;
        const int row = (int)quad * 32 + lane;               // row inside the query block == TMEM lane
        const uint32_t t_lane = (quad * 32u) << 16;
        const uint32_t p_row = p_s + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
        uint32_t it = 0, g = 0;
        for (; iter.valid(); iter.next(), it++) {
            int b, h, qb, n_blocks;
            iter.decode(b, h, qb, n_blocks);
            const int q0 = qb * BQ, q_idx = q0 + row, row0 = b * p.seq_len;
            const uint32_t obuf = it & 1u;
            const uint32_t tmem_o = tmem_base + 256 + obuf * 128;
            float m = -INFINITY, l = 0.0f;
            for (int j = 0; j < n_blocks; j++, g++) {
                const uint32_t s = g & 1u;
                const uint32_t t_s = tmem_base + t_lane + s * BKEY;
                const int key0 = j * BKEY;
                const bool diag = key0 + BKEY - 1 > q0;           // block touches the causal boundary (or the sequence end)
                mbar_wait(bar_sf + 8 * s, (g >> 1) & 1u);
                tc_fence_after();
                // ---- the whole score row in registers: four x32 TMEM loads in flight, one wait
                uint32_t sv[BKEY];
#pragma unroll
                for (int c = 0; c < BKEY / 32; c++) tc_ld32_nowait(t_s + c * 32, sv + c * 32);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_se + 8 * s);       // S slot drained: the MMA warp may overwrite it with block g + 2
                // four independent max / sum chains: one softmax warp per scheduler has no other warp to hide a 128-long dependent chain
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < BKEY; i++) {
                    float x = __uint_as_float(sv[i]) * p.scale_log2e;
                    if (diag && (key0 + i > q_idx || key0 + i >= p.seq_len)) x = -INFINITY;
                    sv[i] = __float_as_uint(x);
                    mx4[i & 3] = fmaxf(mx4[i & 3], x);
                }
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                // Lazy running maximum: the reference point only moves when the row maximum grew by more than 2^8 (exp2 of anything
                // below m + 8 is < 256: no overflow in fp32 or bf16, and the normalisation by l cancels the offset exactly), so the O
                // accumulator in TMEM is almost never rescaled after the first block
                float m_new = m;
                if (mx > m + 8.0f) m_new = mx;
                const float m_use = m_new == -INFINITY ? 0.0f : m_new;   // fully masked row (query beyond the sequence): keep exp2 finite
                const float alpha = ex2(m - m_use);                       // m = -inf on the first block -> 0; unchanged reference -> 1
                // ---- O rescale (needs O += P V of the previous block retired) and the P buffer free
                if (j > 0) {
                    mbar_wait(bar_pfree, (g - 1) & 1u);
                    tc_fence_after();
                    if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                        for (int c = 0; c < D / 32; c++) {
                            uint32_t o[32];
                            tc_ld32(tmem_o + t_lane + c * 32, o);
#pragma unroll
                            for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tc_st32(tmem_o + t_lane + c * 32, o);
                        }
                    }
                }
                // ---- probabilities -> bf16 A tile (SWIZZLE_128B, K-major over the keys), row sum
                float sum4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int c = 0; c < BKEY / 32; c++) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float e0 = ex2(__uint_as_float(sv[c * 32 + 2 * i]) - m_use), e1 = ex2(__uint_as_float(sv[c * 32 + 2 * i + 1]) - m_use);
                        pk[i] = cvt_bf16x2(e1, e0);
                        // the row sum uses the ROUNDED probabilities, so that O / l normalises exactly what P V accumulated
                        sum4[(2 * i) & 3] += __uint_as_float(pk[i] << 16);
                        sum4[(2 * i + 1) & 3] += __uint_as_float(pk[i] & 0xffff0000u);
                    }
                    // 32 keys = 4 chunks of 16 bytes; chunk index inside the 64-key sub-tile, XOR-swizzled with the row
                    const uint32_t sub = (uint32_t)(c >> 1) * kSubBytes;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t chunk = (uint32_t)((c & 1) * 4 + q);
                        const uint32_t addr = p_row + sub + ((chunk ^ (uint32_t)(row & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                    }
                }
                l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
                m = m_new;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy P stores -> visible to the tensor core's async proxy
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pf);   // P ready, O rescaled
            }
            // ---- epilogue: O / l -> bf16 (the MMA warp is already on the next item's S)
            mbar_wait(bar_pfree, (g - 1) & 1u);
            tc_fence_after();
            const float inv = l > 0.0f ? 1.0f / l : 0.0f;
            uint16_t* orow = p.out + ((int64_t)(row0 + q_idx) * p.n_heads + h) * D;
            uint32_t ov[D];
#pragma unroll
            for (int c = 0; c < D / 32; c++) tc_ld32_nowait(tmem_o + t_lane + c * 32, ov + c * 32);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_oe + 8 * obuf);      // this O buffer may be overwritten by item it + 2
            if (q_idx < p.seq_len) {
#pragma unroll
                for (int q = 0; q < D / 8; q++) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        w[i] = cvt_bf16x2(__uint_as_float(ov[8 * q + 2 * i + 1]) * inv, __uint_as_float(ov[8 * q + 2 * i]) * inv);
                    *reinterpret_cast<uint4*>(orow + q * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                }
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

// ------------------------------------------------------------------------------------------------ attention core, two streams per SM
// The persistent kernel above runs ONE softmax warpgroup per SM: every barrier hop of its chain (TMEM load, P store + proxy fence,
// MMA issue, commit) is exposed -- ncu: issue slots 34 % busy, tensor pipe 19 %, 455 us per call vs cuDNN's 216 us.  This variant puts
// TWO independent copies of that pipeline on an SM (producer warp + MMA warp + 4 softmax warps each, 12 warps), each walking its own
// item stream with its own Q / K / V^T / P buffers, S slots and O accumulator, with 64-key blocks so that both fit: per stream
// Q 32 KB + 2 x (K 16 KB + V^T 16 KB) + P 16 KB = 112 KB of shared memory and S 2 x 64 + O 128 = 256 TMEM columns.  While one
// stream waits on a hop the other computes; tcgen05.mma / commit are issued by one thread per stream on disjoint accumulators.
constexpr int BK2 = 64;
constexpr int kThreads2 = 384;

template <int D>
__global__ void __launch_bounds__(kThreads2, 1)
attn_core2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_vt,
                  const AttnParams p) {
    constexpr int NSUB = D / SUBK;
    constexpr int kQBytes = NSUB * kSubBytes;               // Q tile [128 x D]
    constexpr int kKSub = BK2 * SUBK * 2;                   // K sub-tile [64 keys x 64] = 8 KB
    constexpr int kKBytes = NSUB * kKSub;                   // K tile [64 keys x D]
    constexpr int kVBytes = D * SUBK * 2;                   // V^T tile [D x 64 keys]
    constexpr int kPBytes = kSubBytes;                      // P tile [128 x 64 keys]
    constexpr int kStageBytes = kKBytes + kVBytes;
    constexpr int kStreamBytes = kQBytes + 2 * kStageBytes + kPBytes;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = warp / 6, role = warp - 6 * w;            // stream, role inside the stream (0 producer, 1 MMA, 2..5 softmax)
    const uint32_t base = smem_u32(smem) + (uint32_t)w * kStreamBytes;
    const uint32_t q_s = base, kv_s = base + kQBytes, p_s = kv_s + 2 * kStageBytes;
    const uint32_t bars = smem_u32(smem) + 2 * kStreamBytes + (uint32_t)w * 128;
    // per stream: q_full, q_empty, kv_full[2], kv_empty[2], s_full[2], s_empty[2], o_empty, p_full, p_free
    const uint32_t bar_qf = bars, bar_qe = bars + 8, bar_kvf = bars + 16, bar_kve = bars + 32, bar_sf = bars + 48, bar_se = bars + 64,
                   bar_oe = bars + 80, bar_pf = bars + 88, bar_pfree = bars + 96;
    uint32_t* tmem_slot = (uint32_t*)(smem + 2 * kStreamBytes + 256);

    if (lane == 0 && role == 0) {
        mbar_init(bar_qf, 1); mbar_init(bar_qe, 1);
        for (int s = 0; s < 2; s++) { mbar_init(bar_kvf + 8 * s, 1); mbar_init(bar_kve + 8 * s, 1); mbar_init(bar_sf + 8 * s, 1); mbar_init(bar_se + 8 * s, 4); }
        mbar_init(bar_oe, 4); mbar_init(bar_pf, 4); mbar_init(bar_pfree, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot + (uint32_t)w * 256;   // this stream's 256 columns: S slots [0, 64) / [64, 128), O [128, 128 + D)
    const uint32_t tmem_o = tmem_base + 128;

    ItemIter iter;
    iter.t = 2 * (int)blockIdx.x + w; iter.stride = 2 * (int)gridDim.x;
    iter.heads_total = p.samples * p.n_heads; iter.end = iter.heads_total * p.q_blocks;
    iter.q_blocks = p.q_blocks; iter.n_heads = p.n_heads; iter.n_kv = p.n_kv; iter.seq_len = p.seq_len;
    const int kv_rep = p.n_heads / p.n_kv;
    const int key_blocks_total = (p.seq_len + BK2 - 1) / BK2;
    auto blocks_of = [&](int qb) { return min(2 * (qb + 1), key_blocks_total); };   // 64-key blocks up to the end of the 128-query block

    if (role == 0) {
        // ------------------------------------------------------------------ TMA producer of this stream
        if (lane == 0) {
            uint32_t it = 0, g = 0;
            for (; iter.valid(); iter.next(), it++) {
                int b, h, qb, nb_unused;
                iter.decode(b, h, qb, nb_unused);
                const int nb = blocks_of(qb);
                const int hk = h / kv_rep, row0 = b * p.seq_len;
                mbar_wait(bar_qe, (it & 1u) ^ 1u);                            // the S MMAs of the previous item have retired
                mbar_arrive_expect_tx(bar_qf, kQBytes);
#pragma unroll
                for (int t = 0; t < NSUB; t++) tma_load_2d(q_s + t * kSubBytes, &map_q, h * D + t * SUBK, row0 + qb * BQ, bar_qf);
                for (int j = 0; j < nb; j++, g++) {
                    const uint32_t s = g & 1u;
                    mbar_wait(bar_kve + 8 * s, ((g >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_kvf + 8 * s, kStageBytes);
                    const uint32_t k_dst = kv_s + s * kStageBytes, v_dst = k_dst + kKBytes;
#pragma unroll
                    for (int t = 0; t < NSUB; t++)
                        tma_load_2d(k_dst + t * kKSub, &map_k, (p.n_heads + hk) * D + t * SUBK, row0 + j * BK2, bar_kvf + 8 * s);
                    tma_load_2d(v_dst, &map_vt, j * BK2, (b * p.n_kv + hk) * D, bar_kvf + 8 * s);
                }
            }
        }
    } else if (role == 1) {
        // ------------------------------------------------------------------ MMA issuer of this stream
        if (lane == 0) {
            ItemIter ia = iter, ib = iter;
            uint32_t it_a = 0, g_a = 0, it_b = 0, g_b = 0;
            int ja = 0, nba = 0;
            bool a_open = false;
            auto advance_a = [&]() -> bool {
                if (!a_open) {
                    if (!ia.valid()) return false;
                    int b, h, qb, nb_unused;
                    ia.decode(b, h, qb, nb_unused);
                    nba = blocks_of(qb);
                    ja = 0;
                    a_open = true;
                    mbar_wait(bar_qf, it_a & 1u);
                }
                const uint32_t s = g_a & 1u;
                mbar_wait(bar_kvf + 8 * s, (g_a >> 1) & 1u);
                mbar_wait(bar_se + 8 * s, ((g_a >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t k_addr = kv_s + s * kStageBytes;
#pragma unroll
                for (int k = 0; k < D / 16; k++) {
                    const uint64_t ad = make_desc(q_s + (k / 4) * kSubBytes) + 2 * (k % 4);
                    const uint64_t bd = make_desc(k_addr + (k / 4) * kKSub) + 2 * (k % 4);
                    tc_mma(tmem_base + s * BK2, ad, bd, idesc_n(BK2), k ? 1u : 0u);
                }
                tc_commit(bar_sf + 8 * s);
                g_a++;
                if (++ja == nba) {
                    tc_commit(bar_qe);
                    a_open = false;
                    ia.next();
                    it_a++;
                }
                return true;
            };
            advance_a();
            for (; ib.valid(); ib.next(), it_b++) {
                int b, h, qb, nb_unused;
                ib.decode(b, h, qb, nb_unused);
                const int nbb = blocks_of(qb);
                for (int jb = 0; jb < nbb; jb++, g_b++) {
                    // exactly ONE block of look-ahead: S of block g_b + 1 (the next item's first block at an item boundary) before
                    // P V of block g_b.  More would need the ring stage that block g_b still holds (its P V frees it) -- a
                    // two-block look-ahead at item boundaries dead-locked the first version of this kernel on multi-item streams.
                    advance_a();
                    const uint32_t s = g_b & 1u;
                    mbar_wait(bar_pf, g_b & 1u);
                    if (jb == 0) mbar_wait(bar_oe, (it_b & 1u) ^ 1u);   // the previous item's epilogue has read O
                    tc_fence_after();
                    const uint32_t v_addr = kv_s + s * kStageBytes + kKBytes;
#pragma unroll
                    for (int k = 0; k < BK2 / 16; k++)
                        tc_mma(tmem_o, make_desc(p_s) + 2 * k, make_desc(v_addr) + 2 * k, idesc_n(D), (jb | k) ? 1u : 0u);
                    tc_commit(bar_kve + 8 * s);
                    tc_commit(bar_pfree);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax + epilogue of this stream: one query row per thread
        const uint32_t quad = (uint32_t)warp & 3u;
        const int row = (int)quad * 32 + lane;
        const uint32_t t_lane = (quad * 32u) << 16;
        const uint32_t p_row = p_s + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
        uint32_t it = 0, g = 0;
        for (; iter.valid(); iter.next(), it++) {
            int b, h, qb, nb_unused;
            iter.decode(b, h, qb, nb_unused);
            const int n_blocks = blocks_of(qb);
            const int q0 = qb * BQ, q_idx = q0 + row, row0 = b * p.seq_len;
            float m = -INFINITY, l = 0.0f;
            for (int j = 0; j < n_blocks; j++, g++) {
                const uint32_t s = g & 1u;
                const uint32_t t_s = tmem_base + t_lane + s * BK2;
                const int key0 = j * BK2;
                const bool diag = key0 + BK2 - 1 > q0;
                mbar_wait(bar_sf + 8 * s, (g >> 1) & 1u);
                tc_fence_after();
                uint32_t sv[BK2];
                tc_ld32_nowait(t_s, sv);
                tc_ld32_nowait(t_s + 32, sv + 32);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_se + 8 * s);
                // Lean softmax (ncu on the first version: ~14 issued instructions per score; one softmax warp per scheduler and
                // stream cannot afford that): the maximum is taken on the RAW scores (the scale is positive), the scale and the
                // reference point fold into one FFMA in front of ex2, the causal / tail mask only runs on the blocks that need it,
                // and the row sum adds the fp32 probabilities (4 chains): max + FFMA + MUFU + 1/2 cvt + add = 4.5 per score.
                if (diag) {
#pragma unroll
                    for (int i = 0; i < BK2; i++)
                        if (key0 + i > q_idx || key0 + i >= p.seq_len) sv[i] = 0xff800000u;   // -inf
                }
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < BK2; i++) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));   // raw score units
                float m_new = m;
                if (mx > m + p.lazy_raw) m_new = mx;                   // lazy running maximum: 2^8 of head-room (see attn_core_kernel)
                const float mc = (m_new == -INFINITY ? 0.0f : m_new) * p.scale_log2e;
                const float alpha = ex2(m * p.scale_log2e - mc);       // m = -inf on the first block -> 0; unchanged reference -> 1
                if (j > 0) {
                    mbar_wait(bar_pfree, (g - 1) & 1u);
                    tc_fence_after();
                    if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                        for (int c = 0; c < D / 32; c++) {
                            uint32_t o[32];
                            tc_ld32(tmem_o + t_lane + c * 32, o);
#pragma unroll
                            for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tc_st32(tmem_o + t_lane + c * 32, o);
                        }
                    }
                }
                float sum4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                uint32_t pk[BK2 / 2];
#pragma unroll
                for (int i = 0; i < BK2 / 2; i++) {
                    const float e0 = ex2(fmaf(__uint_as_float(sv[2 * i]), p.scale_log2e, -mc));
                    const float e1 = ex2(fmaf(__uint_as_float(sv[2 * i + 1]), p.scale_log2e, -mc));
                    pk[i] = cvt_bf16x2(e1, e0);
                    sum4[(2 * i) & 3] += e0;
                    sum4[(2 * i + 1) & 3] += e1;
                }
#pragma unroll
                for (int q = 0; q < BK2 / 8; q++) {                    // 8 chunks of 16 bytes, XOR-swizzled with the row
                    const uint32_t addr = p_row + (((uint32_t)q ^ (uint32_t)(row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                }
                l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
                m = m_new;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pf);
            }
            // ---- epilogue: O / l -> bf16, in two halves (register budget of a 12-warp CTA)
            mbar_wait(bar_pfree, (g - 1) & 1u);
            tc_fence_after();
            const float inv = l > 0.0f ? 1.0f / l : 0.0f;
            uint16_t* orow = p.out + ((int64_t)(row0 + q_idx) * p.n_heads + h) * D;
#pragma unroll
            for (int hf = 0; hf < D / 64; hf++) {
                uint32_t ov[64];
                tc_ld32_nowait(tmem_o + t_lane + hf * 64, ov);
                tc_ld32_nowait(tmem_o + t_lane + hf * 64 + 32, ov + 32);
                tc_wait_ld();
                if (hf == D / 64 - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_oe);               // O may be overwritten by the next item
                }
                if (q_idx < p.seq_len) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        uint32_t wv[4];
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            wv[i] = cvt_bf16x2(__uint_as_float(ov[8 * q + 2 * i + 1]) * inv, __uint_as_float(ov[8 * q + 2 * i]) * inv);
                        *reinterpret_cast<uint4*>(orow + hf * 64 + q * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(kTmemCols) : "memory");
    }
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

// 2-D bf16 tensor [rows, cols] with a row pitch, box = [box_rows x 64 columns], SWIZZLE_128B
int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows) {
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return B200Q_ECUDA; }
    const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows}, strides[1] = {(uint64_t)pitch_elems * 2};
    const uint32_t box[2] = {SUBK, (uint32_t)box_rows}, estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return B200Q_ECUDA; }
    return B200Q_OK;
}

template <int D>
int launch_attn(const CUtensorMap& mqk, const CUtensorMap& mvt, const AttnParams& p, cudaStream_t st) {
    // 2 Q buffers + 2 x (K + V^T) stages + P + barriers / TMEM slot + 1 KB alignment slack: 225.25 KB for d = 128 (limit 227 KB)
    constexpr size_t smem = 2 * (size_t)(D / SUBK) * kSubBytes + 2 * ((size_t)(D / SUBK) * kSubBytes + 2 * (D * SUBK * 2)) + 2 * kSubBytes + 256 + 1024;
    cudaFuncSetAttribute(attn_core_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t grid = min((int64_t)p.samples * p.n_heads * p.q_blocks, (int64_t)kNumSMs);
    attn_core_kernel<D><<<(unsigned)grid, kThreads, smem, st>>>(mqk, mvt, p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

template <int D>
int launch_attn2(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mvt, const AttnParams& p, cudaStream_t st) {
    constexpr size_t stream = (size_t)(D / SUBK) * kSubBytes + 2 * ((size_t)(D / SUBK) * (BK2 * SUBK * 2) + (size_t)D * SUBK * 2) + kSubBytes;
    constexpr size_t smem = 2 * stream + 512 + 1024;
    cudaFuncSetAttribute(attn_core2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t items = (int64_t)p.samples * p.n_heads * p.q_blocks;
    const int64_t grid = min((items + 1) / 2, (int64_t)kNumSMs);
    attn_core2_kernel<D><<<(unsigned)grid, kThreads2, smem, st>>>(mq, mk, mvt, p);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

}  // namespace
}  // namespace b200q

using namespace b200q;

extern "C" {

int64_t b200q_attention_workspace(int64_t tokens, int32_t n_kv, int32_t head_dim, int32_t seq_len) {
    if (seq_len <= 0) return 0;
    const int64_t s_pad = ((int64_t)seq_len + 63) / 64 * 64;
    return (tokens / seq_len) * n_kv * head_dim * s_pad * 2;
}

int b200q_attention_core(const void* qkv, int64_t tokens, int32_t n_heads, int32_t n_kv, int32_t head_dim, int32_t seq_len, void* out,
                         void* workspace, int64_t workspace_bytes, void* stream) {
    B200Q_REQUIRE(qkv && out && workspace, "b200q_attention_core: NULL pointer");
    B200Q_REQUIRE(head_dim == 64 || head_dim == 128, "head_dim must be 64 or 128, got %d", (int)head_dim);
    B200Q_REQUIRE(n_heads >= 1 && n_kv >= 1 && n_heads % n_kv == 0, "n_heads must be a multiple of n_kv");
    B200Q_REQUIRE(seq_len >= 1 && tokens >= seq_len && tokens % seq_len == 0, "tokens must be whole samples of seq_len");
    B200Q_REQUIRE(workspace_bytes >= b200q_attention_workspace(tokens, n_kv, head_dim, seq_len), "attention workspace too small");
    B200Q_REQUIRE((((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)workspace) & 15) == 0, "operands must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t row_elems = (int64_t)(n_heads + 2 * n_kv) * head_dim;
    const int samples = (int)(tokens / seq_len);
    const int s_pad = (seq_len + 63) / 64 * 64;
    B200Q_REQUIRE((int64_t)samples * n_heads * ((seq_len + BQ - 1) / BQ) < (1ll << 31), "too many attention work items");
    // V^T per (sample, kv head)
    {
        dim3 grid((unsigned)((s_pad + 31) / 32), (unsigned)(head_dim / 32), (unsigned)(samples * n_kv));
        B200Q_REQUIRE(grid.z <= 65535, "samples * n_kv out of range for one launch");
        attn_transpose_v_kernel<<<grid, dim3(32, 8), 0, st>>>((const uint16_t*)qkv, row_elems, (n_heads + n_kv) * head_dim, head_dim, seq_len, s_pad,
                                                              n_kv, (uint16_t*)workspace);
        B200Q_CHECK_LAUNCH();
    }
    CUtensorMap mqk, mk64, mvt;
    if (int rc = make_map(&mqk, qkv, tokens, row_elems, row_elems, 128)) return rc;
    if (int rc = make_map(&mk64, qkv, tokens, row_elems, row_elems, BK2)) return rc;
    if (int rc = make_map(&mvt, workspace, (int64_t)samples * n_kv * head_dim, s_pad, s_pad, head_dim)) return rc;
    AttnParams p;
    p.n_heads = n_heads; p.n_kv = n_kv; p.seq_len = seq_len; p.q_blocks = (seq_len + BQ - 1) / BQ; p.samples = samples;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)head_dim);
    p.lazy_raw = 8.0f / p.scale_log2e;
    p.out = (uint16_t*)out;
    // default: two pipelines per SM on 64-key blocks; B200Q_ATTN_ONE_STREAM=1 selects the single-stream kernel (128-key blocks) for A/B
    static const bool one_stream = getenv("B200Q_ATTN_ONE_STREAM") != nullptr;
    if (one_stream) return head_dim == 128 ? launch_attn<128>(mqk, mvt, p, st) : launch_attn<64>(mqk, mvt, p, st);
    return head_dim == 128 ? launch_attn2<128>(mqk, mk64, mvt, p, st) : launch_attn2<64>(mqk, mk64, mvt, p, st);
}

}  // extern "C"
