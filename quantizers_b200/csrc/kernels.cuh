// kernels.cuh -- internal launcher interface between the .cu translation units and abi.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200q {

struct TileParams;

enum : int { MODE_COMPRESS = 0, MODE_QUANT = 1, MODE_FQ = 2, MODE_OBS_FQ = 3, MODE_QUANT_PACK = 4 };

// ---- GROUP / TENSOR_GROUP (quant_group.cu)
struct GroupParams {
    const void* w;
    int64_t rows, cols;
    int32_t group, nbits, symmetric, has_zp;
    // qparams: outputs in MODE_COMPRESS, inputs in MODE_QUANT / MODE_FQ
    void* scale;            // T [b, rows, G]  (FP4 compress: e4m3 bytes)
    const int8_t* zp_in;    // int8 [b, rows, G] or null
    int32_t* zp_packed;     // int32 [b, ceil(rows/pf), G]  (compress, asym)
    int8_t* zp_scratch = nullptr;  // optional int8 [b, rows, G] workspace (TMA kernel: plain stores + a row-pack kernel instead of atomics)
    const float* gs;        // fp32 [b] (FP4) or null
    int32_t gs_stride;      // 1: per batch entry, 0: shared
    const float* col_scale; // fp32 [cols]: AWQ smoothing scale (MODE_OBS_FQ) or null
    int64_t col_scale_stride; // 0: one scale vector; cols: one per batch entry applied to the SAME weight (AWQ ratio grid)
    int64_t out_batch_stride; // AWQ ratio grid: elements between consecutive ratios' outputs
    void* out;              // codes / packed words / T values, depending on mode
};
template <int MODE> int dispatch_group(int dt, int qt, const GroupParams& p, int64_t batch, cudaStream_t st);
// bf16 fast paths (issue-budget tuned); return B200Q_ENOSYS when the scheme/shape is not covered
int launch_group_tma(int qt, const GroupParams& p, int64_t batch, cudaStream_t st);
int launch_group_tma_supplied(int qt, const GroupParams& p, int64_t batch, cudaStream_t st);  // caller's qparams (quantize_pack)
int launch_awq_fq_grid_fast(const GroupParams& p, int n_ratios, cudaStream_t st);  // bf16 INT4 AWQ per-ratio weight update (awq_fq_fast.cu)
bool tma_paths_enabled();
int launch_group_fast(int qt, const GroupParams& p, int64_t batch, cudaStream_t st);
int launch_block_fp8_fast(const TileParams& p, int64_t batch, cudaStream_t st);
int launch_tensor_fp8_fast(const TileParams& p, int64_t batch, cudaStream_t st);
int launch_nvfp4_fast(const GroupParams& p, int64_t batch, cudaStream_t st);
int launch_nvfp4_supplied(const GroupParams& p, int64_t batch, cudaStream_t st, int op = 0);  // caller's bf16 group scales + global scale; op 0 pack, 1 quantize (values), 2 fake_quantize
int launch_nvfp4_fused(const GroupParams& p, int64_t batch, int span, float* gs_out, uint32_t* sync, cudaStream_t st,
                       uint16_t* loc_scratch = nullptr);  // loc_scratch: optional bf16 [batch * rows * cols / 16]
int64_t nvfp4_resident_workspace(int64_t batch, int64_t rows, int64_t cols);  // bytes of sync words the persistent kernel needs
int launch_nvfp4_resident(const GroupParams& p, int64_t batch, int span, float* gs_out, uint32_t* sync, cudaStream_t st);
bool fast_paths_enabled();

// ---- generic element-wise path with caller-supplied qparams, any strategy (quant_elementwise.cu)
struct ElemParams {
    const void* x;          // T values, or codes for dequantize
    int64_t rows, cols;
    int32_t strategy, group, bh, bw, nbits, has_zp;
    const void* scale;      // T, qparam grid
    const int8_t* zp;       // int8 or null
    const float* gs;        // fp32[1] or null
    void* out;
};
enum : int { EW_QUANT = 0, EW_FQ = 1, EW_DEQUANT = 2 };
int launch_elementwise(int op, int dt, int qt, const ElemParams& p, cudaStream_t st);
// bf16 fast path (elementwise_fast.cu): quantize / fake_quantize, INT4 or FP8, qparam constant per aligned 8 elements; B200Q_ENOSYS otherwise
int launch_elementwise_fast(int op, int qt, const ElemParams& p, cudaStream_t st);
int launch_dequant_int8_fast(const ElemParams& p, cudaStream_t st);

// ---- observers / qparams (observers.cu)
int launch_minmax(int dt, const void* w, int64_t batch, int64_t rows, int64_t cols, int strategy, int group, int bh, int bw,
                  void* mn, void* mx, cudaStream_t st);
int launch_global_scale(int dt, const void* x, int64_t batch, int64_t numel, float* state, int running, float* gs,
                        cudaStream_t st);
int launch_span_min(float* gs, int64_t batch, int span, cudaStream_t st);
int launch_qparams(int dt, int qt, int nbits, int symmetric, const void* mn, const void* mx, int64_t n, const float* gs,
                   void* scale, int8_t* zp, cudaStream_t st);

// O4: workspace = uint32 [n_chunks + batch]
int launch_mse_minmax(int dt, int qt, int nbits, int symmetric, int strategy, int group, const void* w, int64_t batch, int64_t rows,
                      int64_t cols, const float* gs, float maxshrink, int patience, int grid_n, float norm, void* mn, void* mx,
                      uint32_t* workspace, cudaStream_t st);

// ---- pack / unpack (pack.cu)
int launch_pack_int32(const int8_t* v, int64_t rows, int64_t cols, int nbits, int packed_dim, int32_t* out, cudaStream_t st);
int launch_unpack_int32(const int32_t* p, int64_t rows, int64_t cols, int nbits, int packed_dim, int8_t* out, cudaStream_t st);
int launch_pack_fp4(int dt, const void* x, int64_t rows, int64_t cols, uint8_t* out, cudaStream_t st);
int launch_unpack_fp4(int dt, const uint8_t* p, int64_t rows, int64_t cols, void* out, cudaStream_t st);

int launch_decompress_int_packed(int dt, const int32_t* packed, const void* scale, const int32_t* zp_packed, int64_t batch, int64_t rows,
                                 int64_t cols, int group, int nbits, void* out, cudaStream_t st);
int launch_decompress_nvfp4(int dt, const uint8_t* packed, const uint8_t* scale, const float* gs, int gs_stride, int64_t batch, int64_t rows,
                            int64_t cols, void* out, cudaStream_t st);

// ---- fused CHANNEL / BLOCK / TENSOR compress (quant_tile.cu)
struct TileParams {
    const void* w;
    int64_t rows, cols;
    int32_t nbits, symmetric, has_zp;
    void* scale;          // T
    int32_t* zp_packed;   // INT asym channel
    void* out;            // packed int32 (INT) / e4m3 bytes (FP8)
    float* workspace;     // TENSOR: fp32 absmax bits per batch
};
int launch_channel_compress(int dt, int qt, const TileParams& p, int64_t batch, cudaStream_t st);
// bf16 decode at HBM speed (decode_fast.cu); B200Q_ENOSYS when the scheme / shape is not covered
int launch_decode_int4_fast(const int32_t* packed, const void* scale, const int32_t* zp_packed, int64_t batch, int64_t rows, int64_t cols, int group,
                            void* out, cudaStream_t st);
int launch_decode_fp8_fast(const uint8_t* codes, int64_t rows, int64_t cols, int strategy, int group, int bh, int bw, const void* scale, void* out,
                           cudaStream_t st);
int launch_decode_nvfp4_fast(const uint8_t* packed, const uint8_t* scale, const float* gs, int gs_stride, int64_t batch, int64_t rows, int64_t cols,
                             void* out, cudaStream_t st);
// bf16 issue-tuned CHANNEL compress (quant_channel_fast.cu): FP8 and INT4; B200Q_ENOSYS when the shape is not covered
int launch_channel_fast(int qt, const TileParams& p, int64_t batch, cudaStream_t st);
int launch_block_fp8_compress(int dt, const TileParams& p, int64_t batch, cudaStream_t st);
int launch_tensor_fp8_compress(int dt, const TileParams& p, int64_t batch, cudaStream_t st);

// ---- AWQ statistics (awq_stats.cu)
int launch_abs_sum_cols(int dt, const void* x, int64_t tokens, int64_t k, float* acc, cudaStream_t st);
int launch_wmean(int dt, const void* w, int64_t rows, int64_t cols, int group, double* acc, cudaStream_t st);
int launch_awq_scales(const float* x_mean, const float* w_mean, int64_t k, const float* ratios, int n_ratios, int duo,
                      float* scales, cudaStream_t st);
int launch_moe_combine(const void* y, const int32_t* row, const void* w, int64_t tokens, int top_k, int64_t h, const void* init, void* out,
                       cudaStream_t st);
int launch_sq_err(int dt, const void* a, const void* b, int64_t n, float* acc, cudaStream_t st);

}  // namespace b200q
