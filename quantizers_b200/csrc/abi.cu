// abi.cu -- the extern "C" surface declared in include/b200q.h.  Argument validation mirrors the reference's
// ValueErrors: the Python shims turn a negative return code + b200q_last_error() into ValueError.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "../../include/b200q.h"
#include "common.cuh"
#include "kernels.cuh"

namespace b200q {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool tma_paths_enabled() {  // B200Q_DISABLE_TMA=1 selects the register-tile fast kernels instead
    const char* e = getenv("B200Q_DISABLE_TMA");
    return !(e && e[0] == '1');
}
bool fast_paths_enabled() {  // B200Q_DISABLE_FAST=1 forces the generic kernels (A/B parity tests)
    const char* e = getenv("B200Q_DISABLE_FAST");
    return !(e && e[0] == '1');
}
}  // namespace b200q

using namespace b200q;

#define REQ_PTR(p) B200Q_REQUIRE((p) != nullptr, #p " must not be NULL")

static int check_scheme(const b200q_scheme* s) {
    REQ_PTR(s);
    B200Q_REQUIRE(s->dtype >= 0 && s->dtype <= 2, "unsupported dtype %d", s->dtype);
    B200Q_REQUIRE(s->qtype >= 0 && s->qtype <= 2, "Invalid quantization type %d", s->qtype);
    if (s->qtype == B200Q_INT) B200Q_REQUIRE(s->num_bits == 4 || s->num_bits == 8, "INT num_bits must be 4 or 8, got %d", s->num_bits);
    if (s->qtype == B200Q_FP8) B200Q_REQUIRE(s->num_bits == 8, "Only num_bits in (4, 8) are supported");
    if (s->qtype == B200Q_FP4) B200Q_REQUIRE(s->num_bits == 4, "Only num_bits in (4, 8) are supported");
    if (s->qtype != B200Q_INT) B200Q_REQUIRE(s->symmetric, "Asymmetric Quantization is not supported for float types");
    B200Q_REQUIRE(s->strategy >= 0 && s->strategy <= 3, "unknown strategy %d", s->strategy);
    return 0;
}

extern "C" {

const char* b200q_last_error(void) { return g_err; }
int b200q_version(void) { return 100; }

int64_t b200q_compress_int_workspace(int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* sc) {
    if (sc == nullptr || sc->qtype != B200Q_INT || sc->symmetric || sc->strategy != B200Q_GROUP || sc->group_size <= 0) return 0;
    return batch * rows * (cols / sc->group_size);  // one int8 zero point per group
}

int b200q_compress_int_packed(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* sc,
                              int32_t* packed, void* scale, int32_t* zp_packed, void* stream) {
    return b200q_compress_int_packed_ws(weight, batch, rows, cols, sc, packed, scale, zp_packed, nullptr, 0, stream);
}

int b200q_compress_int_packed_ws(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* sc,
                                 int32_t* packed, void* scale, int32_t* zp_packed, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    B200Q_REQUIRE(sc->qtype == B200Q_INT, "pack-quantized needs an INT scheme");
    REQ_PTR(weight); REQ_PTR(packed); REQ_PTR(scale);
    if (!sc->symmetric) B200Q_REQUIRE(zp_packed != nullptr, "Asymmetric quant requires zero-point values");
    cudaStream_t st = (cudaStream_t)stream;
    if (sc->strategy == B200Q_GROUP) {
        GroupParams p{};
        p.w = weight; p.rows = rows; p.cols = cols; p.group = sc->group_size; p.nbits = sc->num_bits;
        p.symmetric = sc->symmetric; p.has_zp = 1; p.scale = scale; p.zp_packed = zp_packed; p.out = packed;
        if (workspace != nullptr && !sc->symmetric && workspace_bytes >= b200q_compress_int_workspace(batch, rows, cols, sc) &&
            getenv("B200Q_ZP_ATOMIC") == nullptr)
            p.zp_scratch = (int8_t*)workspace;
        if (sc->dtype == B200Q_BF16 && fast_paths_enabled()) {
            int rc = tma_paths_enabled() ? launch_group_tma(QT_INT, p, batch, st) : B200Q_ENOSYS;
            if (rc == B200Q_ENOSYS) rc = launch_group_fast(QT_INT, p, batch, st);
            if (rc != B200Q_ENOSYS) return rc;
        }
        return dispatch_group<MODE_COMPRESS>(sc->dtype, QT_INT, p, batch, st);
    }
    if (sc->strategy == B200Q_CHANNEL) {
        TileParams p{};
        p.w = weight; p.rows = rows; p.cols = cols; p.nbits = sc->num_bits; p.symmetric = sc->symmetric; p.has_zp = 1;
        p.scale = scale; p.zp_packed = zp_packed; p.out = packed;
        if (sc->dtype == B200Q_BF16 && fast_paths_enabled()) {
            const int rc = launch_channel_fast(QT_INT, p, batch, st);
            if (rc != B200Q_ENOSYS) return rc;
        }
        return launch_channel_compress(sc->dtype, QT_INT, p, batch, st);
    }
    set_error("pack-quantized fused compress supports GROUP and CHANNEL strategies (got %d)", sc->strategy);
    return B200Q_EINVAL;
}

int b200q_compress_fp8(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* sc, uint8_t* q,
                       void* scale, void* workspace, void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    B200Q_REQUIRE(sc->qtype == B200Q_FP8, "float-quantized needs an FP8 scheme");
    REQ_PTR(weight); REQ_PTR(q); REQ_PTR(scale);
    cudaStream_t st = (cudaStream_t)stream;
    if (sc->strategy == B200Q_GROUP) {
        GroupParams p{};
        p.w = weight; p.rows = rows; p.cols = cols; p.group = sc->group_size; p.nbits = 8; p.symmetric = 1;
        p.has_zp = sc->has_zp; p.scale = scale; p.out = q;
        if (sc->dtype == B200Q_BF16 && fast_paths_enabled()) {
            int rc = tma_paths_enabled() ? launch_group_tma(QT_FP8, p, batch, st) : B200Q_ENOSYS;
            if (rc == B200Q_ENOSYS) rc = launch_group_fast(QT_FP8, p, batch, st);
            if (rc != B200Q_ENOSYS) return rc;
        }
        return dispatch_group<MODE_COMPRESS>(sc->dtype, QT_FP8, p, batch, st);
    }
    TileParams p{};
    p.w = weight; p.rows = rows; p.cols = cols; p.nbits = 8; p.symmetric = 1; p.has_zp = sc->has_zp; p.scale = scale;
    p.out = q; p.workspace = (float*)workspace;
    if (sc->strategy == B200Q_CHANNEL) {
        if (sc->dtype == B200Q_BF16 && fast_paths_enabled()) {
            const int rc = launch_channel_fast(QT_FP8, p, batch, st);
            if (rc != B200Q_ENOSYS) return rc;
        }
        return launch_channel_compress(sc->dtype, QT_FP8, p, batch, st);
    }
    if (sc->strategy == B200Q_BLOCK) {
        B200Q_REQUIRE(sc->block_h == 128 && sc->block_w == 128, "fused block compress supports block_structure [128,128], got [%d,%d]",
                      sc->block_h, sc->block_w);
        if (sc->dtype == B200Q_BF16 && fast_paths_enabled()) {
            const int rc = launch_block_fp8_fast(p, batch, st);
            if (rc != B200Q_ENOSYS) return rc;
        }
        return launch_block_fp8_compress(sc->dtype, p, batch, st);
    }
    if (sc->dtype == B200Q_BF16 && fast_paths_enabled()) {
        const int rc = launch_tensor_fp8_fast(p, batch, st);
        if (rc != B200Q_ENOSYS) return rc;
    }
    return launch_tensor_fp8_compress(sc->dtype, p, batch, st);
}

static int compress_nvfp4_impl(const void* weight, int64_t batch, int64_t rows, int64_t cols, int32_t dtype, int32_t compute_global,
                               int32_t fuse_span, float* global_scale, uint8_t* packed, uint8_t* scale_e4m3, void* workspace,
                               int64_t workspace_bytes, cudaStream_t st) {
    REQ_PTR(weight); REQ_PTR(global_scale); REQ_PTR(packed); REQ_PTR(scale_e4m3);
    B200Q_REQUIRE(cols % 16 == 0, "tensor column shape must be divisible by the given group_size 16 but got %lld", (long long)cols);
    B200Q_REQUIRE(fuse_span >= 1 && batch % fuse_span == 0, "fuse_span must divide the batch");
    if (rows * cols == 0 || batch == 0) return B200Q_OK;
    GroupParams p{};
    p.w = weight; p.rows = rows; p.cols = cols; p.group = 16; p.nbits = 4; p.symmetric = 1; p.has_zp = 1;
    p.scale = scale_e4m3; p.gs = global_scale; p.gs_stride = 1; p.out = packed;
    const bool fast = dtype == B200Q_BF16 && fast_paths_enabled();
    if (compute_global) {
        int rc = B200Q_ENOSYS;
        if (fast && workspace && workspace_bytes >= 8 * (batch / fuse_span) && (((uintptr_t)workspace) & 3) == 0)
        {
            // one launch, one HBM read (the compress pass re-reads through the L2).  Two schedulings of the same idea: one short
            // CTA per 32-128 KB item (default), or a persistent warp-specialised CTA per SM (B200Q_FP4_PERSISTENT=1, needs one
            // sync word per tile in the workspace).  Both are bound by the ALU pipe at ~0.55 of the HBM roofline.
            const char* v = getenv("B200Q_FP4_PERSISTENT");
            if (v && v[0] == '1' && workspace_bytes >= nvfp4_resident_workspace(batch, rows, cols))
                rc = launch_nvfp4_resident(p, batch, fuse_span, global_scale, (uint32_t*)workspace, st);
            if (rc == B200Q_ENOSYS) {
                // workspace layout: [8 bytes per span: sync words][pad to 256][2 bytes per group: T(|max| / 6) scratch, optional]
                const int64_t sync_bytes = (8 * (batch / fuse_span) + 255) / 256 * 256;
                const int64_t loc_bytes = 2 * batch * rows * (cols / 16);
                // measured (bench moe_nvfp4, 9.66 GB of experts): 2 905 GB/s WITH the scratch vs 3 044 without -- the extra stores keep the
                // |max| CTAs resident longer than the ALU work they save in the compress pass; opt-in only (B200Q_FP4_LOC=1)
                static const bool use_loc = getenv("B200Q_FP4_LOC") != nullptr;
                uint16_t* loc = use_loc && workspace_bytes >= sync_bytes + loc_bytes ? (uint16_t*)((char*)workspace + sync_bytes) : nullptr;
                rc = launch_nvfp4_fused(p, batch, fuse_span, global_scale, (uint32_t*)workspace, st, loc);
            }
        }
        if (rc != B200Q_ENOSYS) return rc;
        // generic: min/max state staged in the first 8*batch bytes of the (not yet written) scale output buffer
        B200Q_REQUIRE(rows * cols / 16 >= 8, "weight too small to stage the min/max state");
        float* state = (float*)scale_e4m3;
        B200Q_REQUIRE(((uintptr_t)state & 3) == 0, "scale buffer must be 4-byte aligned");
        if (int rc2 = launch_global_scale(dtype, weight, batch, rows * cols, state, 0, global_scale, st)) return rc2;
        if (fuse_span > 1) {
            if (int rc2 = launch_span_min(global_scale, batch, fuse_span, st)) return rc2;
        }
    }
    if (fast) {
        int rc = tma_paths_enabled() ? launch_group_tma(QT_FP4, p, batch, st) : B200Q_ENOSYS;
        if (rc == B200Q_ENOSYS) rc = launch_nvfp4_fast(p, batch, st);
        if (rc != B200Q_ENOSYS) return rc;
    }
    return dispatch_group<MODE_COMPRESS>(dtype, QT_FP4, p, batch, st);
}

int b200q_compress_nvfp4(const void* weight, int64_t batch, int64_t rows, int64_t cols, int32_t dtype, int32_t compute_global,
                         float* global_scale, uint8_t* packed, uint8_t* scale_e4m3, void* stream) {
    return compress_nvfp4_impl(weight, batch, rows, cols, dtype, compute_global, 1, global_scale, packed, scale_e4m3, nullptr, 0,
                               (cudaStream_t)stream);
}

int b200q_compress_nvfp4_fused(const void* weight, int64_t batch, int64_t rows, int64_t cols, int32_t dtype, int32_t fuse_span,
                               float* global_scale, uint8_t* packed, uint8_t* scale_e4m3, void* workspace, int64_t workspace_bytes,
                               void* stream) {
    return compress_nvfp4_impl(weight, batch, rows, cols, dtype, 1, fuse_span, global_scale, packed, scale_e4m3, workspace, workspace_bytes,
                               (cudaStream_t)stream);
}

int64_t b200q_compress_nvfp4_workspace(int64_t batch, int64_t rows, int64_t cols, int32_t fuse_span) {
    const int64_t spans = fuse_span > 0 ? (batch + fuse_span - 1) / fuse_span : batch;
    // sync words (padded) + the per-group T(|max| / 6) scratch of the fused kernel; the persistent variant needs its own sync words
    static const bool use_loc = getenv("B200Q_FP4_LOC") != nullptr;
    const int64_t a = (8 * spans + 255) / 256 * 256 + (use_loc ? 2 * batch * rows * (cols / 16) : 0), b = nvfp4_resident_workspace(batch, rows, cols);
    return a > b ? a : b;
}

int b200q_minmax(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* sc, void* mn, void* mx,
                 void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    REQ_PTR(weight); REQ_PTR(mn); REQ_PTR(mx);
    return launch_minmax(sc->dtype, weight, batch, rows, cols, sc->strategy, sc->group_size, sc->block_h, sc->block_w, mn, mx,
                         (cudaStream_t)stream);
}

int b200q_mse_minmax(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* sc, const float* global_scale,
                     float maxshrink, int32_t patience, int32_t grid, float norm, void* mn, void* mx, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    REQ_PTR(weight); REQ_PTR(mn); REQ_PTR(mx); REQ_PTR(workspace);
    B200Q_REQUIRE(sc->strategy == B200Q_GROUP || sc->strategy == B200Q_CHANNEL, "mse observer: GROUP / TENSOR_GROUP / CHANNEL strategies only");
    const int64_t len = sc->strategy == B200Q_CHANNEL ? cols : sc->group_size;
    B200Q_REQUIRE(len > 0 && cols % len == 0, "tensor column shape must be divisible by the given group_size %lld but got %lld",
                  (long long)len, (long long)cols);
    const int64_t n_chunks = batch * rows * (cols / len);
    B200Q_REQUIRE(workspace_bytes >= 4 * (n_chunks + batch) && (((uintptr_t)workspace) & 3) == 0,
                  "mse observer: workspace of %lld bytes needed", (long long)(4 * (n_chunks + batch)));
    return launch_mse_minmax(sc->dtype, sc->qtype, sc->num_bits, sc->symmetric, sc->strategy, sc->group_size, weight, batch, rows, cols,
                             global_scale, maxshrink, patience, grid, norm, mn, mx, (uint32_t*)workspace, (cudaStream_t)stream);
}

int b200q_global_scale(const void* x, int64_t batch, int64_t numel, int32_t dtype, float* minmax_state, int32_t running,
                       float* global_scale, void* stream) {
    REQ_PTR(x); REQ_PTR(minmax_state);
    return launch_global_scale(dtype, x, batch, numel, minmax_state, running, global_scale, (cudaStream_t)stream);
}

int b200q_calculate_qparams(const void* mn, const void* mx, int64_t n, const b200q_scheme* sc, const float* global_scale,
                            void* scale, int8_t* zp, void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    REQ_PTR(mn); REQ_PTR(mx); REQ_PTR(scale);
    return launch_qparams(sc->dtype, sc->qtype, sc->num_bits, sc->symmetric, mn, mx, n, global_scale, scale, zp,
                          (cudaStream_t)stream);
}

static int elementwise(int op, const void* x, int64_t rows, int64_t cols, const b200q_scheme* sc, const void* scale,
                       const int8_t* zp, const float* gs, void* out, void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    REQ_PTR(x); REQ_PTR(scale); REQ_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec_ok = cols % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0;
    const int g = sc->group_size;
    static const bool ew_fast = getenv("B200Q_ELEMENTWISE_LEGACY") == nullptr;  // A/B switch
    if (ew_fast && op != EW_DEQUANT && sc->dtype == B200Q_BF16 && gs == nullptr && vec_ok && fast_paths_enabled()) {
        ElemParams f{};
        f.x = x; f.rows = rows; f.cols = cols; f.strategy = sc->strategy; f.group = g; f.bh = sc->block_h; f.bw = sc->block_w;
        f.nbits = sc->num_bits; f.has_zp = sc->has_zp; f.scale = scale; f.zp = zp; f.gs = gs; f.out = out;
        const int rc = launch_elementwise_fast(op, sc->qtype, f, st);
        if (rc != B200Q_ENOSYS) return rc;
    }
    if (ew_fast && op != EW_DEQUANT && sc->qtype == B200Q_FP4 && sc->dtype == B200Q_BF16 && sc->strategy == B200Q_GROUP && g == 16 && gs != nullptr &&
        vec_ok && cols % 16 == 0 && fast_paths_enabled()) {
        GroupParams p{};
        p.w = x; p.rows = rows; p.cols = cols; p.group = 16; p.nbits = 4; p.symmetric = 1; p.has_zp = sc->has_zp;
        p.scale = const_cast<void*>(scale); p.gs = gs; p.gs_stride = 0; p.out = out;
        const int rc = launch_nvfp4_supplied(p, 1, st, op == EW_QUANT ? 1 : 2);
        if (rc != B200Q_ENOSYS) return rc;
    }
    if (op != EW_DEQUANT && sc->strategy == B200Q_GROUP && vec_ok && (g == 16 || g == 32 || g == 64 || g == 128 || g == 256) &&
        cols % g == 0 && (sc->qtype != B200Q_FP4 || gs != nullptr)) {
        GroupParams p{};
        p.w = x; p.rows = rows; p.cols = cols; p.group = g; p.nbits = sc->num_bits; p.symmetric = sc->symmetric;
        p.has_zp = sc->has_zp; p.scale = const_cast<void*>(scale); p.zp_in = zp; p.gs = gs; p.gs_stride = 0; p.out = out;
        return op == EW_QUANT ? dispatch_group<MODE_QUANT>(sc->dtype, sc->qtype, p, 1, st)
                              : dispatch_group<MODE_FQ>(sc->dtype, sc->qtype, p, 1, st);
    }
    if (ew_fast && op == EW_DEQUANT && sc->qtype == B200Q_INT && sc->dtype == B200Q_BF16 && gs == nullptr && fast_paths_enabled()) {
        ElemParams f{};
        f.x = x; f.rows = rows; f.cols = cols; f.strategy = sc->strategy; f.group = g; f.bh = sc->block_h; f.bw = sc->block_w;
        f.nbits = sc->num_bits; f.has_zp = sc->has_zp; f.scale = scale; f.zp = zp; f.gs = gs; f.out = out;
        const int rc = launch_dequant_int8_fast(f, st);
        if (rc != B200Q_ENOSYS) return rc;
    }
    if (op == EW_DEQUANT && sc->qtype == B200Q_FP8 && sc->dtype == B200Q_BF16 && zp == nullptr && gs == nullptr && fast_paths_enabled()) {
        const int rc = launch_decode_fp8_fast((const uint8_t*)x, rows, cols, sc->strategy, g, sc->block_h, sc->block_w, scale, out, st);
        if (rc != B200Q_ENOSYS) return rc;
    }
    ElemParams p{};
    p.x = x; p.rows = rows; p.cols = cols; p.strategy = sc->strategy; p.group = g; p.bh = sc->block_h; p.bw = sc->block_w;
    p.nbits = sc->num_bits; p.has_zp = sc->has_zp; p.scale = scale; p.zp = zp; p.gs = gs; p.out = out;
    return launch_elementwise(op, sc->dtype, sc->qtype, p, st);
}
int b200q_quantize(const void* x, int64_t rows, int64_t cols, const b200q_scheme* sc, const void* scale, const int8_t* zp,
                   const float* gs, uint8_t* codes, void* stream) {
    return elementwise(EW_QUANT, x, rows, cols, sc, scale, zp, gs, codes, stream);
}
int b200q_quantize_pack(const void* x, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* sc, const void* scale,
                        const int8_t* zp, const float* gs, void* packed, void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    REQ_PTR(x); REQ_PTR(scale); REQ_PTR(packed);
    B200Q_REQUIRE(sc->strategy == B200Q_GROUP, "b200q_quantize_pack handles GROUP strategies; use b200q_quantize + b200q_pack_int32 otherwise");
    if (sc->qtype == B200Q_FP4) REQ_PTR(gs);
    GroupParams p{};
    p.w = x; p.rows = rows; p.cols = cols; p.group = sc->group_size; p.nbits = sc->num_bits; p.symmetric = sc->symmetric;
    p.has_zp = sc->has_zp; p.scale = const_cast<void*>(scale); p.zp_in = zp; p.gs = gs; p.gs_stride = 0; p.out = packed;
    // row-walk kernel (elementwise_fast.cu; op 3 = quantize + nibble pack, FP8 "pack" is plain quantize) or the TMA group kernel:
    // B200Q_QPACK=rows|tma overrides the per-format default
    static const int qpack_mode = [] { const char* v = getenv("B200Q_QPACK"); return !v ? 0 : (v[0] == 'r' ? 1 : 2); }();
    // measured (398 MB matrix, fraction of the HBM roofline, rows / tma): INT4 g128 sym 0.85 / 0.80, g128 asym 0.78 / 0.86, g32 sym 0.86 / 0.71,
    // FP8 g32 0.92 / 0.69
    const bool rows_first = qpack_mode == 1 || (qpack_mode == 0 && (sc->qtype == B200Q_FP8 || sc->group_size < 128 || zp == nullptr));
    if (rows_first && sc->dtype == B200Q_BF16 && fast_paths_enabled() && (sc->qtype == B200Q_FP8 || (sc->qtype == B200Q_INT && sc->num_bits == 4))) {
        ElemParams f{};
        f.x = x; f.rows = batch * rows; f.cols = cols; f.strategy = B200Q_GROUP; f.group = sc->group_size; f.nbits = sc->num_bits;
        f.has_zp = sc->has_zp; f.scale = scale; f.zp = zp; f.out = packed;
        const int rc = launch_elementwise_fast(sc->qtype == B200Q_FP8 ? EW_QUANT : 3, sc->qtype, f, (cudaStream_t)stream);
        if (rc != B200Q_ENOSYS) return rc;
    }
    if (sc->dtype == B200Q_BF16 && fast_paths_enabled() && tma_paths_enabled() && (sc->qtype == B200Q_FP8 || (sc->qtype == B200Q_INT && sc->num_bits == 4))) {
        const int rc = launch_group_tma_supplied(sc->qtype == B200Q_FP8 ? QT_FP8 : QT_INT, p, batch, (cudaStream_t)stream);
        if (rc != B200Q_ENOSYS) return rc;
    }
    if (sc->dtype == B200Q_BF16 && fast_paths_enabled() && sc->qtype == B200Q_FP4 && sc->group_size == 16) {
        const int rc = launch_nvfp4_supplied(p, batch, (cudaStream_t)stream);
        if (rc != B200Q_ENOSYS) return rc;
    }
    return dispatch_group<MODE_QUANT_PACK>(sc->dtype, sc->qtype, p, batch, (cudaStream_t)stream);
}
int b200q_fake_quantize(const void* x, int64_t rows, int64_t cols, const b200q_scheme* sc, const void* scale, const int8_t* zp,
                        const float* gs, void* out, void* stream) {
    return elementwise(EW_FQ, x, rows, cols, sc, scale, zp, gs, out, stream);
}
int b200q_dequantize(const void* codes, int64_t rows, int64_t cols, const b200q_scheme* sc, const void* scale, const int8_t* zp,
                     const float* gs, void* out, void* stream) {
    return elementwise(EW_DEQUANT, codes, rows, cols, sc, scale, zp, gs, out, stream);
}

int b200q_decompress_int_packed(const int32_t* packed, const void* scale, const int32_t* zp_packed, int64_t batch, int64_t rows, int64_t cols,
                                const b200q_scheme* sc, void* out, void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    REQ_PTR(packed); REQ_PTR(scale); REQ_PTR(out);
    B200Q_REQUIRE(sc->qtype == B200Q_INT && (sc->strategy == B200Q_GROUP || sc->strategy == B200Q_CHANNEL), "pack-quantized decompress: INT, group or channel");
    if (sc->dtype == B200Q_BF16 && sc->num_bits == 4 && fast_paths_enabled()) {
        const int rc = launch_decode_int4_fast(packed, scale, zp_packed, batch, rows, cols, sc->strategy == B200Q_GROUP ? sc->group_size : 0, out,
                                               (cudaStream_t)stream);
        if (rc != B200Q_ENOSYS) return rc;
    }
    return launch_decompress_int_packed(sc->dtype, packed, scale, zp_packed, batch, rows, cols, sc->strategy == B200Q_GROUP ? sc->group_size : 0,
                                        sc->num_bits, out, (cudaStream_t)stream);
}
int b200q_decompress_nvfp4(const uint8_t* packed, const uint8_t* scale_e4m3, const float* global_scale, int64_t batch, int64_t rows, int64_t cols,
                           int32_t dtype, void* out, void* stream) {
    REQ_PTR(packed); REQ_PTR(scale_e4m3); REQ_PTR(global_scale); REQ_PTR(out);
    if (dtype == B200Q_BF16 && fast_paths_enabled()) {
        const int rc = launch_decode_nvfp4_fast(packed, scale_e4m3, global_scale, 1, batch, rows, cols, out, (cudaStream_t)stream);
        if (rc != B200Q_ENOSYS) return rc;
    }
    return launch_decompress_nvfp4(dtype, packed, scale_e4m3, global_scale, 1, batch, rows, cols, out, (cudaStream_t)stream);
}
int b200q_pack_int32(const int8_t* value, int64_t rows, int64_t cols, int32_t num_bits, int32_t packed_dim, int32_t* packed,
                     void* stream) {
    REQ_PTR(value); REQ_PTR(packed);
    return launch_pack_int32(value, rows, cols, num_bits, packed_dim, packed, (cudaStream_t)stream);
}
int b200q_unpack_int32(const int32_t* packed, int64_t rows, int64_t cols, int32_t num_bits, int32_t packed_dim, int8_t* value,
                       void* stream) {
    REQ_PTR(value); REQ_PTR(packed);
    return launch_unpack_int32(packed, rows, cols, num_bits, packed_dim, value, (cudaStream_t)stream);
}
int b200q_pack_fp4(const void* x, int64_t rows, int64_t cols, int32_t dtype, uint8_t* packed, void* stream) {
    REQ_PTR(x); REQ_PTR(packed);
    return launch_pack_fp4(dtype, x, rows, cols, packed, (cudaStream_t)stream);
}
int b200q_unpack_fp4(const uint8_t* packed, int64_t rows, int64_t cols, int32_t dtype, void* out, void* stream) {
    REQ_PTR(out); REQ_PTR(packed);
    return launch_unpack_fp4(dtype, packed, rows, cols, out, (cudaStream_t)stream);
}

int b200q_abs_sum_cols(const void* x, int64_t tokens, int64_t k, int32_t dtype, float* acc, void* stream) {
    REQ_PTR(acc);
    if (tokens == 0) return 0;
    REQ_PTR(x);
    return launch_abs_sum_cols(dtype, x, tokens, k, acc, (cudaStream_t)stream);
}
int b200q_wmean_accumulate(const void* weight, int64_t rows, int64_t cols, int32_t dtype, int32_t group_size, double* acc,
                           void* stream) {
    REQ_PTR(weight); REQ_PTR(acc);
    return launch_wmean(dtype, weight, rows, cols, group_size, acc, (cudaStream_t)stream);
}
int b200q_awq_scales(const float* x_mean, const float* w_mean, int64_t k, const float* ratios_host, int32_t n_ratios,
                     int32_t duo, float* scales, void* stream) {
    REQ_PTR(x_mean); REQ_PTR(ratios_host); REQ_PTR(scales);
    B200Q_REQUIRE(n_ratios >= 0 && n_ratios <= 256, "n_ratios out of range");
    return launch_awq_scales(x_mean, w_mean, k, ratios_host, n_ratios, duo, scales, (cudaStream_t)stream);
}
int b200q_awq_scaled_fake_quantize(const void* weight, int64_t rows, int64_t cols, const b200q_scheme* sc, const float* scales,
                                   void* out, void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    REQ_PTR(weight); REQ_PTR(out);
    B200Q_REQUIRE(sc->strategy == B200Q_GROUP, "AWQ fused fake-quantize supports GROUP strategies");
    B200Q_REQUIRE(sc->qtype != B200Q_FP4, "AWQ over NVFP4 goes through b200q_global_scale + b200q_fake_quantize");
    GroupParams p{};
    p.w = weight; p.rows = rows; p.cols = cols; p.group = sc->group_size; p.nbits = sc->num_bits; p.symmetric = sc->symmetric;
    p.has_zp = sc->has_zp; p.col_scale = scales; p.out = out;
    if (sc->dtype == B200Q_BF16 && sc->qtype == B200Q_INT && scales != nullptr && fast_paths_enabled()) {
        const int rc = launch_awq_fq_grid_fast(p, 1, (cudaStream_t)stream);
        if (rc != B200Q_ENOSYS) return rc;
    }
    return dispatch_group<MODE_OBS_FQ>(sc->dtype, sc->qtype, p, 1, (cudaStream_t)stream);
}
int b200q_awq_scaled_fake_quantize_grid(const void* weight, int64_t rows, int64_t cols, const b200q_scheme* sc, const float* scales,
                                        int32_t n_ratios, void* out, int64_t out_stride, void* stream) {
    if (int rc = check_scheme(sc)) return rc;
    REQ_PTR(weight); REQ_PTR(out); REQ_PTR(scales);
    B200Q_REQUIRE(sc->strategy == B200Q_GROUP, "AWQ fused fake-quantize supports GROUP strategies");
    B200Q_REQUIRE(sc->qtype != B200Q_FP4, "AWQ over NVFP4 goes through b200q_global_scale + b200q_fake_quantize");
    B200Q_REQUIRE(n_ratios >= 1 && cols % 8 == 0, "need n_ratios >= 1 and a multiple of 8 columns");
    B200Q_REQUIRE(out_stride >= rows * cols && out_stride % 8 == 0, "out_stride must cover one weight and keep 16-byte alignment");
    GroupParams p{};
    p.w = weight; p.rows = rows; p.cols = cols; p.group = sc->group_size; p.nbits = sc->num_bits; p.symmetric = sc->symmetric;
    p.has_zp = sc->has_zp; p.col_scale = scales; p.col_scale_stride = cols; p.out_batch_stride = out_stride; p.out = out;
    if (sc->dtype == B200Q_BF16 && sc->qtype == B200Q_INT && fast_paths_enabled()) {
        const int rc = launch_awq_fq_grid_fast(p, n_ratios, (cudaStream_t)stream);
        if (rc != B200Q_ENOSYS) return rc;
    }
    return dispatch_group<MODE_OBS_FQ>(sc->dtype, sc->qtype, p, n_ratios, (cudaStream_t)stream);
}
int b200q_moe_combine(const void* y, const int32_t* row, const void* weight, int64_t tokens, int32_t top_k, int64_t hidden, void* out, void* stream) {
    REQ_PTR(y); REQ_PTR(row); REQ_PTR(weight); REQ_PTR(out);
    return launch_moe_combine(y, row, weight, tokens, top_k, hidden, nullptr, out, (cudaStream_t)stream);
}
int b200q_moe_combine_acc(const void* y, const int32_t* row, const void* weight, int64_t tokens, int32_t top_k, int64_t hidden, const void* init,
                          void* out, void* stream) {
    REQ_PTR(y); REQ_PTR(row); REQ_PTR(weight); REQ_PTR(out);
    return launch_moe_combine(y, row, weight, tokens, top_k, hidden, init, out, (cudaStream_t)stream);
}
int b200q_sq_err_accumulate(const void* y_ref, const void* y_q, int64_t numel, int32_t dtype, float* acc, void* stream) {
    REQ_PTR(y_ref); REQ_PTR(y_q); REQ_PTR(acc);
    return launch_sq_err(dtype, y_ref, y_q, numel, acc, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------- host pipeline
struct b200q_pipeline {
    static constexpr int NS = 3;
    int device;
    int64_t cap;  // bytes per slot for the weight
    int64_t codes_cap, scale_cap, zp_cap, ws_cap;  // bytes of the per-slot output / workspace buffers (every job is checked against them)
    cudaStream_t s_in, s_run, s_out;
    void* d_w[NS];
    void* d_codes[NS];
    void* d_scale[NS];
    void* d_zp[NS];
    float* d_gs[NS];
    void* d_ws[NS];
    cudaEvent_t ev_in[NS], ev_run[NS], ev_out[NS];
    bool used[NS];
    int next;
};

#define CU(x)                                                                              \
    do {                                                                                   \
        cudaError_t e__ = (x);                                                             \
        if (e__ != cudaSuccess) { set_error("%s: %s", #x, cudaGetErrorString(e__)); return B200Q_ECUDA; } \
    } while (0)

int b200q_pipeline_create(b200q_pipeline** out, int64_t max_weight_bytes, int32_t device) {
    REQ_PTR(out);
    B200Q_REQUIRE(max_weight_bytes > 0, "max_weight_bytes must be positive");
    CU(cudaSetDevice(device));
    b200q_pipeline* p = new b200q_pipeline();
    p->device = device; p->cap = max_weight_bytes; p->next = 0;
    p->codes_cap = max_weight_bytes / 2 + 256;      // <= 1 byte per 2-byte element
    p->scale_cap = max_weight_bytes / 16 + 4096;    // <= T per 16 elements
    p->zp_cap = max_weight_bytes / 64 + 4096;
    p->ws_cap = max_weight_bytes / 32 + 65536 * (int64_t)sizeof(float);  // int8 zero-point scratch (one per group >= 16) / TENSOR |max| words
    CU(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < b200q_pipeline::NS; i++) {
        CU(cudaMalloc(&p->d_w[i], max_weight_bytes));
        CU(cudaMalloc(&p->d_codes[i], p->codes_cap));
        CU(cudaMalloc(&p->d_scale[i], p->scale_cap));
        CU(cudaMalloc(&p->d_zp[i], p->zp_cap));
        CU(cudaMalloc((void**)&p->d_gs[i], 65536 * sizeof(float)));
        CU(cudaMalloc(&p->d_ws[i], p->ws_cap));
        CU(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&p->ev_run[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
        p->used[i] = false;
    }
    *out = p;
    return 0;
}
int b200q_pipeline_destroy(b200q_pipeline* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->s_in); cudaStreamSynchronize(p->s_run); cudaStreamSynchronize(p->s_out);
    for (int i = 0; i < b200q_pipeline::NS; i++) {
        cudaFree(p->d_w[i]); cudaFree(p->d_codes[i]); cudaFree(p->d_scale[i]); cudaFree(p->d_zp[i]); cudaFree(p->d_gs[i]);
        cudaFree(p->d_ws[i]);
        cudaEventDestroy(p->ev_in[i]); cudaEventDestroy(p->ev_run[i]); cudaEventDestroy(p->ev_out[i]);
    }
    cudaStreamDestroy(p->s_in); cudaStreamDestroy(p->s_run); cudaStreamDestroy(p->s_out);
    delete p;
    return 0;
}
int b200q_pipeline_sync(b200q_pipeline* p) {
    REQ_PTR(p);
    CU(cudaStreamSynchronize(p->s_in));
    CU(cudaStreamSynchronize(p->s_run));
    CU(cudaStreamSynchronize(p->s_out));
    return 0;
}

int b200q_pipeline_compress_host(b200q_pipeline* p, const void* weight_host, int64_t batch, int64_t rows, int64_t cols,
                                 const b200q_scheme* sc, void* codes_host, void* scale_host, void* zp_host,
                                 float* global_scale_host) {
    REQ_PTR(p); REQ_PTR(weight_host); REQ_PTR(codes_host); REQ_PTR(scale_host);
    if (int rc = check_scheme(sc)) return rc;
    const int64_t esz = sc->dtype == B200Q_F32 ? 4 : 2;
    const int64_t n = batch * rows * cols, wbytes = n * esz;
    B200Q_REQUIRE(wbytes <= p->cap, "weight of %lld bytes exceeds the pipeline slot (%lld)", (long long)wbytes, (long long)p->cap);
    B200Q_REQUIRE(batch <= 65535, "batch out of range");
    int64_t qrows, qcols;
    switch (sc->strategy) {
    case B200Q_TENSOR: qrows = 1; qcols = 1; break;
    case B200Q_CHANNEL: qrows = rows; qcols = 1; break;
    case B200Q_GROUP: B200Q_REQUIRE(sc->group_size > 0, "group_size required"); qrows = rows; qcols = cols / sc->group_size; break;
    default: qrows = (rows + 127) / 128; qcols = (cols + 127) / 128; break;
    }
    int64_t code_bytes, scale_bytes, zp_bytes = 0;
    if (sc->qtype == B200Q_INT) {
        const int pf = 32 / sc->num_bits;
        code_bytes = batch * rows * ((cols + pf - 1) / pf) * 4;
        scale_bytes = batch * qrows * qcols * esz;
        if (!sc->symmetric) { REQ_PTR(zp_host); zp_bytes = batch * ((qrows + pf - 1) / pf) * qcols * 4; }
    } else if (sc->qtype == B200Q_FP8) {
        code_bytes = n; scale_bytes = batch * qrows * qcols * esz;
    } else {
        REQ_PTR(global_scale_host);
        code_bytes = n / 2; scale_bytes = n / 16;
    }
    // the slot buffers are sized from the weight bytes for the common schemes; a CHANNEL scheme on a narrow weight or 8-bit
    // packing with padded columns needs more than that -- refuse instead of overflowing on the device and in the D2H copy
    B200Q_REQUIRE(code_bytes <= p->codes_cap, "packed codes of %lld bytes exceed the pipeline slot (%lld): create the pipeline with a larger max_weight_bytes",
                  (long long)code_bytes, (long long)p->codes_cap);
    B200Q_REQUIRE(scale_bytes <= p->scale_cap, "scales of %lld bytes exceed the pipeline slot (%lld): create the pipeline with a larger max_weight_bytes",
                  (long long)scale_bytes, (long long)p->scale_cap);
    B200Q_REQUIRE(zp_bytes <= p->zp_cap, "zero points of %lld bytes exceed the pipeline slot (%lld): create the pipeline with a larger max_weight_bytes",
                  (long long)zp_bytes, (long long)p->zp_cap);
    CU(cudaSetDevice(p->device));
    const int s = p->next;
    p->next = (p->next + 1) % b200q_pipeline::NS;
    if (p->used[s]) CU(cudaStreamWaitEvent(p->s_in, p->ev_out[s], 0));  // slot reuse: previous D2H must have drained
    CU(cudaMemcpyAsync(p->d_w[s], weight_host, wbytes, cudaMemcpyHostToDevice, p->s_in));
    CU(cudaEventRecord(p->ev_in[s], p->s_in));
    CU(cudaStreamWaitEvent(p->s_run, p->ev_in[s], 0));
    int rc;
    if (sc->qtype == B200Q_INT)
        rc = b200q_compress_int_packed_ws(p->d_w[s], batch, rows, cols, sc, (int32_t*)p->d_codes[s], p->d_scale[s], (int32_t*)p->d_zp[s],
                                          p->d_ws[s], p->ws_cap, p->s_run);
    else if (sc->qtype == B200Q_FP8)
        rc = b200q_compress_fp8(p->d_w[s], batch, rows, cols, sc, (uint8_t*)p->d_codes[s], p->d_scale[s], p->d_ws[s], p->s_run);
    else
        rc = b200q_compress_nvfp4(p->d_w[s], batch, rows, cols, sc->dtype, 1, p->d_gs[s], (uint8_t*)p->d_codes[s], (uint8_t*)p->d_scale[s], p->s_run);
    if (rc) return rc;
    CU(cudaEventRecord(p->ev_run[s], p->s_run));
    CU(cudaStreamWaitEvent(p->s_out, p->ev_run[s], 0));
    CU(cudaMemcpyAsync(codes_host, p->d_codes[s], code_bytes, cudaMemcpyDeviceToHost, p->s_out));
    CU(cudaMemcpyAsync(scale_host, p->d_scale[s], scale_bytes, cudaMemcpyDeviceToHost, p->s_out));
    if (zp_bytes) CU(cudaMemcpyAsync(zp_host, p->d_zp[s], zp_bytes, cudaMemcpyDeviceToHost, p->s_out));
    if (sc->qtype == B200Q_FP4) CU(cudaMemcpyAsync(global_scale_host, p->d_gs[s], batch * sizeof(float), cudaMemcpyDeviceToHost, p->s_out));
    CU(cudaEventRecord(p->ev_out[s], p->s_out));
    p->used[s] = true;
    return 0;
}

}  // extern "C"
