/*
 * b200q.h -- C ABI of libb200q.so: the B200 (sm_100a) implementation of the quantization hot path behind
 * mratsim/quantizers' llm-compressor / compressed-tensors recipes.
 *
 * The reference has no native code and no FFI; its seams are Python callables.  Every entry
 * point below names the reference callable it replaces (CT: = compressed_tensors 0.15.0.1, the reference's
 * pinned dependency, /root/reference/pyproject.toml:8; LLMC: = llmcompressor >= 0.9, pyproject.toml:9).
 * The Python binding a maintainer would add is quantizers_b200/_lib.py (ctypes); see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless the name ends in _host.
 *   - the caller owns every buffer; the library allocates nothing persistent except inside an explicit
 *     b200q_pipeline handle.
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it, nothing synchronises unless
 *     stated.  No global mutable state: thread-safe when each thread uses its own stream.
 *   - return 0 on success, negative errno-style code otherwise; b200q_last_error() gives the message
 *     (thread-local).
 *   - weights are row-major [batch, rows, cols] (batch = experts of a MoE layer, 1 for a dense Linear).
 */
#ifndef B200Q_H
#define B200Q_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* tensor dtypes (weight / scale) */
#define B200Q_BF16 0
#define B200Q_F16 1
#define B200Q_F32 2
/* QuantizationType x num_bits  (CT:quantization/quant_args.py:91-98) */
#define B200Q_INT 0 /* type=int,   num_bits in {4,8} */
#define B200Q_FP8 1 /* type=float, num_bits=8 (float8_e4m3fn) */
#define B200Q_FP4 2 /* type=float, num_bits=4 (e2m1, NVFP4) */
/* QuantizationStrategy (CT:quantization/quant_args.py:100-112) */
#define B200Q_TENSOR 0
#define B200Q_CHANNEL 1
#define B200Q_GROUP 2 /* GROUP and TENSOR_GROUP */
#define B200Q_BLOCK 3

/* mirrors the fields of CT QuantizationArgs that change the arithmetic (quant_args.py:157-408) */
typedef struct b200q_scheme {
    int32_t dtype;      /* B200Q_BF16 / F16 / F32: dtype of the weight, of weight_scale and of all rounding */
    int32_t qtype;      /* B200Q_INT / FP8 / FP4 */
    int32_t num_bits;   /* 4 or 8 */
    int32_t symmetric;  /* 1 / 0 */
    int32_t strategy;   /* B200Q_TENSOR / CHANNEL / GROUP / BLOCK */
    int32_t group_size; /* GROUP: 16, 32, 64, 128 or 256 */
    int32_t block_h;    /* BLOCK: 128 */
    int32_t block_w;    /* BLOCK: 128 */
    int32_t has_zp;     /* 1 when the caller's state dict carries a zero-point tensor (adds +0: -0.0 -> +0.0) */
    int32_t reserved;
} b200q_scheme;

const char* b200q_last_error(void);
int b200q_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Fused  observer -> calculate_qparams -> quantize -> pack   (one HBM read of the weight)
 * Replaces, for one Linear weight (or a batch of expert weights):
 *   LLMC update_weight_zp_scale (memoryless_minmax Observer.forward -> CT calculate_qparams, helpers.py:50-137)
 *   + CT Compressor.compress  (pack_quantized/base.py:36-77 | naive_quantized/base.py:34-82 | nvfp4/base.py:40-72)
 * Outputs are exactly the state-dict tensors CT emits.
 * ------------------------------------------------------------------------------------------------------- */

/* pack-quantized (INT4/INT8, GROUP or CHANNEL):
 *   packed     int32 [batch, rows, ceil(cols*num_bits/32)]
 *   scale      T     [batch, rows, n_groups]            (n_groups = cols/group_size, 1 for CHANNEL)
 *   zp_packed  int32 [batch, ceil(rows*num_bits/32), n_groups]  (asymmetric only; may be NULL when symmetric) */
int b200q_compress_int_packed(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* scheme,
                              int32_t* packed, void* scale, int32_t* zp_packed, void* stream);
/* The same call with a caller-owned device workspace of b200q_compress_int_workspace(...) bytes.  Asymmetric GROUP schemes then
 * write their zero points with plain int8 stores and row-pack them (pack_to_int32(zp, packed_dim=0),
 * CT:compressors/pack_quantized/base.py:70-73) in a second small kernel, instead of one atomic per group into a zeroed buffer.
 * workspace may be NULL (or too small): the call then behaves exactly like b200q_compress_int_packed.  Identical output bits. */
int64_t b200q_compress_int_workspace(int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* scheme);
int b200q_compress_int_packed_ws(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* scheme,
                                 int32_t* packed, void* scale, int32_t* zp_packed, void* workspace, int64_t workspace_bytes,
                                 void* stream);

/* float-quantized (FP8 e4m3; CHANNEL, GROUP, BLOCK 128x128 or TENSOR):
 *   q      uint8 (float8_e4m3fn bits) [batch, rows, cols]
 *   scale  T [batch, rows, 1] | [batch, rows, n_groups] | [batch, ceil(rows/128), ceil(cols/128)] | [batch, 1]
 *   workspace: TENSOR strategy only, >= 4*batch bytes (device), else may be NULL */
int b200q_compress_fp8(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* scheme,
                       uint8_t* q, void* scale, void* workspace, void* stream);

/* nvfp4-pack-quantized (e2m1 codes, e4m3 per-16 scales, fp32 global scale):
 *   global_scale fp32 [batch] : INPUT when compute_global != 0 is false; otherwise OUTPUT, computed from each
 *                weight's absmax (LLMC Observer.get_global_scale -> CT generate_gparam, helpers.py:309-338).
 *                Fused q/k/v and gate/up siblings: call b200q_global_scale per tensor, take the min on the
 *                host side (LLMC update_fused_layer_weight_global_scales), then pass compute_global = 0.
 *   packed  uint8 [batch, rows, cols/2];  scale_e4m3 uint8 [batch, rows, cols/16] */
int b200q_compress_nvfp4(const void* weight, int64_t batch, int64_t rows, int64_t cols, int32_t dtype,
                         int32_t compute_global, float* global_scale, uint8_t* packed, uint8_t* scale_e4m3, void* stream);

/* Same with the global scale always computed here and shared by `fuse_span` consecutive matrices of the batch
 * (LLMC update_fused_layer_weight_global_scales: gate_proj / up_proj of one expert stacked next to each other get
 * min(gs_gate, gs_up)); global_scale fp32 [batch] is an OUTPUT.  workspace: device scratch (4-byte aligned) of
 * b200q_compress_nvfp4_workspace(...) bytes: with it (bf16) the |max| reduction and the compression are ONE launch and each weight
 * is read from HBM once (the second pass is served by the L2); without a workspace the two-launch path runs.  Identical bits.
 * (B200Q_FP4_LOC=1, experiment: the |max| pass also leaves every group's T(|max| / 6) in a 2-byte-per-group tail of the workspace
 * for the compress pass -- measured slower, see DESIGN.md.) */
int64_t b200q_compress_nvfp4_workspace(int64_t batch, int64_t rows, int64_t cols, int32_t fuse_span);
int b200q_compress_nvfp4_fused(const void* weight, int64_t batch, int64_t rows, int64_t cols, int32_t dtype, int32_t fuse_span,
                               float* global_scale, uint8_t* packed, uint8_t* scale_e4m3, void* workspace, int64_t workspace_bytes,
                               void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Observers  (LLMC observers/min_max.py + observers/base.py, restated in SURVEY.md Appendix A)
 * ------------------------------------------------------------------------------------------------------- */

/* O1: min / max per quantization chunk (flatten_for_calibration + amin/amax).  mn, mx: T, qparam-grid shaped. */
int b200q_minmax(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* scheme, void* mn,
                 void* mx, void* stream);
/* O4: MSE observer (LLMC observers/mse.py `mse`; defaults maxshrink 0.2, patience 5, grid 100, norm 2.4).  For every
 * quantization chunk (GROUP / TENSOR_GROUP: group_size consecutive elements of a row; CHANNEL: a row) the (min, max) range is
 * shrunk by p = 1 - i/grid, i < int(maxshrink * grid), and the range with the smallest sum |fake_quantize(x) - x|^norm is kept,
 * with the reference's tensor-wide early stop (`patience` steps without an improvement in any chunk; per batch entry).
 * mn, mx: T, qparam-grid shaped, feed b200q_calculate_qparams.  global_scale: fp32[1] for NVFP4 (TENSOR_GROUP), else NULL.
 * workspace: device scratch of >= 4 * (number of chunks + batch) bytes.  One HBM read of the weight. */
int b200q_mse_minmax(const void* weight, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* scheme,
                     const float* global_scale, float maxshrink, int32_t patience, int32_t grid, float norm, void* mn, void* mx,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* O2 + Q2: per-tensor global scale = generate_gparam(min, max) -> fp32 [batch].
 * Also O3 (static_minmax activation global scale): pass running != 0 to fold the previous min/max kept in
 * minmax_state (fp32 [batch,2], initialise to {+inf,-inf}) before deriving the scale. */
int b200q_global_scale(const void* x, int64_t batch, int64_t numel, int32_t dtype, float* minmax_state, int32_t running,
                       float* global_scale, void* stream);
/* Q1: calculate_qparams(min_vals, max_vals, args, global_scale) CT:quantization/utils/helpers.py:50-137
 *   scale: T (fp32 when global_scale != NULL; e4m3-rounded values for FP4); zp: int8 (INT only, else untouched) */
int b200q_calculate_qparams(const void* mn, const void* mx, int64_t n, const b200q_scheme* scheme,
                            const float* global_scale, void* scale, int8_t* zp, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Un-fused element-wise path with caller-supplied qparams (AWQ inner loop, decompression, round trips)
 *   scale: T, qparam-grid shaped; zp: int8 (INT) or NULL; global_scale: fp32[1] device or NULL
 * ------------------------------------------------------------------------------------------------------- */
/* Q3: CT quantize()  forward.py:37-73 -> codes: int8 (INT) | e4m3 bytes (FP8) | e2m1 grid values as T (FP4) */
int b200q_quantize(const void* x, int64_t rows, int64_t cols, const b200q_scheme* scheme, const void* scale,
                   const int8_t* zp, const float* global_scale, uint8_t* codes, void* stream);
/* Q3 + Q7/Q8 fused: CT Compressor.compress with the module's existing qparams (compressors/pack_quantized/base.py:36-77,
 * nvfp4/base.py:40-72, naive_quantized/base.py:34-82): codes written directly in storage layout -- int32 words (INT4),
 * offset bytes (INT8), e4m3 bytes (FP8), two e2m1 nibbles per byte (FP4).  GROUP strategies (batch of weights). */
int b200q_quantize_pack(const void* x, int64_t batch, int64_t rows, int64_t cols, const b200q_scheme* scheme,
                        const void* scale, const int8_t* zp, const float* global_scale, void* packed, void* stream);
/* Q4: CT fake_quantize()  forward.py:149-181 -> out: T [rows, cols] */
int b200q_fake_quantize(const void* x, int64_t rows, int64_t cols, const b200q_scheme* scheme, const void* scale,
                        const int8_t* zp, const float* global_scale, void* out, void* stream);
/* Q5: CT dequantize()  forward.py:77-145.  codes: int8 | e4m3 bytes | (FP4) values as T.  out dtype = scheme->dtype */
int b200q_dequantize(const void* codes, int64_t rows, int64_t cols, const b200q_scheme* scheme, const void* scale,
                     const int8_t* zp, const float* global_scale, void* out, void* stream);

/* Compressor.decompress in one pass (CT:compressors/pack_quantized/base.py:79-113, nvfp4/base.py:74-96): packed codes + qparams
 * -> T weights [batch, rows, cols], without the unpacked int8 / T intermediates.  Same layouts as the b200q_compress_* outputs;
 * zp_packed NULL for symmetric schemes (nothing is subtracted, like dequantize(zero_point=None)); global_scale fp32 [batch]. */
int b200q_decompress_int_packed(const int32_t* packed, const void* scale, const int32_t* zp_packed, int64_t batch, int64_t rows, int64_t cols,
                                const b200q_scheme* scheme, void* out, void* stream);
int b200q_decompress_nvfp4(const uint8_t* packed, const uint8_t* scale_e4m3, const float* global_scale, int64_t batch, int64_t rows, int64_t cols,
                           int32_t dtype, void* out, void* stream);

/* Q7/Q9: pack_to_int32 / unpack_from_int32  CT:compressors/pack_quantized/helpers.py:20-161 */
int b200q_pack_int32(const int8_t* value, int64_t rows, int64_t cols, int32_t num_bits, int32_t packed_dim,
                     int32_t* packed, void* stream);
int b200q_unpack_int32(const int32_t* packed, int64_t rows, int64_t cols, int32_t num_bits, int32_t packed_dim,
                       int8_t* value, void* stream);
/* Q8/Q9: pack_fp4_to_uint8 / unpack_fp4_from_uint8  CT:compressors/nvfp4/helpers.py:34-111 */
int b200q_pack_fp4(const void* x, int64_t rows, int64_t cols, int32_t dtype, uint8_t* packed, void* stream);
int b200q_unpack_fp4(const uint8_t* packed, int64_t rows, int64_t cols, int32_t dtype, void* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * AWQ statistics and search  (LLMC modifiers/awq/base.py, restated in SURVEY.md Appendix A)
 * ------------------------------------------------------------------------------------------------------- */
/* O5: _accumulate_mean numerator: acc[k] += sum_t |x[t,k]|  (fp32 [K], caller zero-initialises; all-reduce SUM
 * across token shards, then divide by the token count) */
int b200q_abs_sum_cols(const void* x, int64_t tokens, int64_t k, int32_t dtype, float* acc, void* stream);
/* O5: _compute_layer_means numerator: acc[k] += sum_rows |w| / (group_absmax + 1e-6)  (fp64 [K]) */
int b200q_wmean_accumulate(const void* weight, int64_t rows, int64_t cols, int32_t dtype, int32_t group_size,
                           double* acc, void* stream);
/* W1: scales for one grid point: s = x_mean^r / (w_mean^(1-r) + 1e-4), clamp(min=1e-4)  [duo_scaling]
 *     | x_mean^r clamp(1e-4);   s /= sqrt(max(s) * min(s));  inf/nan -> 1.   fp32 [K] in, fp32 [K] out.
 *     n_ratios rows are produced at once: scales [n_ratios, K], ratios fp32 [n_ratios] (host pointer). */
int b200q_awq_scales(const float* x_mean, const float* w_mean, int64_t k, const float* ratios_host, int32_t n_ratios,
                     int32_t duo_scaling, float* scales, void* stream);
/* W1 inner step, fused: W' = fake_quantize(W * s[None,:]) / s[None,:]  with a fresh memoryless_minmax observer
 * (call_observer + forward_quantize + div), written as T [rows, cols] */
int b200q_awq_scaled_fake_quantize(const void* weight, int64_t rows, int64_t cols, const b200q_scheme* scheme,
                                   const float* scales, void* out, void* stream);
/* The same for all grid points in one launch: scales fp32 [n_ratios, cols] (b200q_awq_scales output); variant r is written at
 * out + r * out_stride elements (out_stride >= rows * cols: the variants of several balance layers interleave in one buffer). */
int b200q_awq_scaled_fake_quantize_grid(const void* weight, int64_t rows, int64_t cols, const b200q_scheme* scheme,
                                        const float* scales, int32_t n_ratios, void* out, int64_t out_stride, void* stream);
/* W3: _compute_loss partial: acc[0] += sum (bf16(y_ref) - bf16(y_q))^2 in fp32 (device fp32 accumulator) */
int b200q_sq_err_accumulate(const void* y_ref, const void* y_q, int64_t numel, int32_t dtype, float* acc, void* stream);
/* W1-W3 fused on the tensor cores: for each of n_ratios pre-fake-quantized weight variants Wq[r] (T [n, k]) compute
 * loss[r] += sum_{t,n} ( bf16(X Wref^T)[t,n] - bf16(X Wq[r]^T)[t,n] )^2  without materialising the outputs.
 * tcgen05 / TMEM bf16 GEMM (fp32 accumulate) with the squared-error reduction in the epilogue.  bf16 only. */
int b200q_awq_gemm_loss(const void* x, int64_t tokens, int64_t k, const void* w_ref, const void* w_q, int64_t n,
                        int32_t n_ratios, float* loss, void* workspace, int64_t workspace_bytes, void* stream);
int64_t b200q_awq_gemm_loss_workspace(int64_t tokens, int64_t k, int64_t n, int32_t n_ratios);
/* Same kernel with either operand varying per ratio (multi-layer parents):
 *   loss[r] += sum ( bf16(A_ref B_ref^T) - bf16(A_r B_r^T) )^2,   A_r = a_q ? a_q[r] : a_ref  (T [n_ratios, tokens, k]),
 *                                                                 B_r = b_q ? b_q[r] : b_ref  (T [n_ratios, n, k]).
 * MLP parent: a_q[r] = silu(x Wg'_r^T) * (x Wu'_r^T), b_ref = W_down; attention parent: a_q[r] = attention output, b_ref = W_o. */
int b200q_awq_gemm_loss_pairs(const void* a_ref, const void* a_q, int64_t tokens, int64_t k, const void* b_ref, const void* b_q,
                              int64_t n, int32_t n_ratios, float* loss, void* workspace, int64_t workspace_bytes, void* stream);
/* W2 first stage of a multi-layer parent, all weight variants in one launch (tcgen05, double-buffered TMEM):
 *   swiglu == 0:  out[v] = bf16(x W[v]^T)                                   W: T [n_variants, n_out, k]      (q/k/v projections)
 *   swiglu != 0:  out[v] = silu(bf16(x Wg[v]^T)) * bf16(x Wu[v]^T)           W: T [n_variants, 2*n_out, k]    (gate rows, then up rows)
 * out: T [n_variants, tokens, n_out], rounded exactly where torch's bf16 Linear / silu / mul round.  bf16 only. */
int b200q_awq_gemm_project(const void* x, int64_t tokens, int64_t k, const void* w, int64_t n_variants, int64_t n_out, int32_t swiglu,
                           void* out, void* stream);

/* W2 of the layer-wide MoE mapping (post_attention_layernorm -> every expert's w1 / w3; parent = the routed sparse-MoE block of
 * transformers' Mixtral / Qwen3-MoE / MiniMax modules): the same projection with one weight PER 128-ROW TILE of x.  The caller
 * sorts the routed (token, expert) pairs expert-major and pads every expert's rows to whole tiles; tile_expert int32
 * [tokens / 128] (device) names the expert of each tile (< 0: padding tile, skipped).  w: T [n_experts, rows, k] with rows = n_out
 * (swiglu == 0) or 2 * n_out (gate rows, then up rows); out: T [tokens, n_out].  One launch replaces n_experts launches. */
int b200q_awq_gemm_project_grouped(const void* x, int64_t tokens, int64_t k, const void* w, int64_t n_experts, int64_t n_out, int32_t swiglu,
                                   const int32_t* tile_expert, void* out, void* stream);
/* ... and its combine step: out[t] = sum_j bf16(y[row[t, j]] * weight[t, j]), accumulated in bf16 with j in ascending expert
 * order (the rounding sequence of the reference's per-expert index_add_); row int32 [tokens, top_k] (< 0: skip), weight T
 * [tokens, top_k], y T [padded rows, hidden], out T [tokens, hidden].  bf16 only. */
int b200q_moe_combine(const void* y, const int32_t* row, const void* weight, int64_t tokens, int32_t top_k, int64_t hidden, void* out,
                      void* stream);
/* ... continuing a running sum: out[t] = init[t] + (this call's rows, same order and rounding); init T [tokens, hidden] may alias out,
 * NULL = zeros.  Expert-parallel ranks own ascending expert ranges, so rank r continues the sequence rank r - 1 left off (its partial
 * output arrives over NCCL send/recv) and the last rank holds the block output bit-identical to the single-GPU one
 * (quantizers_b200/awq.py search_moe_block_mapping_ep). */
int b200q_moe_combine_acc(const void* y, const int32_t* row, const void* weight, int64_t tokens, int32_t top_k, int64_t hidden, const void* init,
                          void* out, void* stream);

/* SURVEY.md §8f rank 4 -- GPTQ Hessian accumulation (LLMC modifiers/gptq: accumulate_hessian; the reference's GPTQ recipes,
 * /root/reference/scripts/old_scripts/main_glm4-gptq.py:108-126): hessian fp32 [features, features] = alpha * hessian + beta * X^T X
 * on the tensor cores (tcgen05, fp32 accumulate, fp32 read-modify-write epilogue).  xt: T = bf16 X^T [features, tokens], contiguous
 * (tokens % 8 == 0).  With n the running sample count: alpha = n / (n + tokens), beta = 2 / (n + tokens). */
int b200q_gptq_hessian_accumulate(const void* xt, int64_t features, int64_t tokens, float alpha, float beta, float* hessian, void* stream);

/* W2, attention parent (input_layernorm -> q/k/v mapping; transformers Qwen3Attention.forward between the projections and
 * SDPA): in-place per-head RMSNorm (weights T [head_dim]) + rotary embedding of the q and k columns of
 * qkv T [tokens, (n_heads + 2 n_kv) * head_dim]; position = token index % seq_len; cos/sin T [seq_len, head_dim].  bf16 only. */
int b200q_qk_norm_rope(void* qkv, int64_t tokens, int32_t n_heads, int32_t n_kv, int32_t head_dim, int32_t seq_len,
                       const void* q_norm_weight, const void* k_norm_weight, const void* cos, const void* sin, float eps, void* stream);

/* W2, attention parent: the causal grouped-query attention core softmax(Q K^T / sqrt(head_dim)) V of every calibration sample
 * (batch-1 forwards of seq_len tokens in the reference's _run_samples; transformers sdpa_attention_forward with is_causal) on the
 * tcgen05 tensor cores: fp32 scores / softmax statistics, bf16 probabilities, fp32 accumulation.
 *   qkv  T [tokens, (n_heads + 2 n_kv) * head_dim]   (q, k normalised + rotated by b200q_qk_norm_rope; read only)
 *   out  T [tokens, n_heads * head_dim]
 *   workspace: b200q_attention_workspace(...) device bytes (V transposed per sample and kv head)
 * tokens = samples * seq_len; head_dim 64 or 128; n_heads % n_kv == 0.  bf16 only. */
int64_t b200q_attention_workspace(int64_t tokens, int32_t n_kv, int32_t head_dim, int32_t seq_len);
int b200q_attention_core(const void* qkv, int64_t tokens, int32_t n_heads, int32_t n_kv, int32_t head_dim, int32_t seq_len, void* out,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Host-buffer pipeline: what LLMC model_free_ptq's per-tensor job does (load -> device -> observe ->
 * compress -> host), /root/reference/scripts/quant_GLM-4.7-Flash-FP8.py:11-24.  Pinned staging buffers and two
 * streams live in the handle; weight_host / outputs are HOST pointers (pinned or pageable).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct b200q_pipeline b200q_pipeline;
int b200q_pipeline_create(b200q_pipeline** out, int64_t max_weight_bytes, int32_t device);
int b200q_pipeline_destroy(b200q_pipeline* p);
/* fmt: 0 pack-quantized, 1 float-quantized, 2 nvfp4-pack-quantized; outputs as in the b200q_compress_* calls.
 * Asynchronous: returns after enqueueing; b200q_pipeline_sync waits for every submitted job. */
int b200q_pipeline_compress_host(b200q_pipeline* p, const void* weight_host, int64_t batch, int64_t rows, int64_t cols,
                                 const b200q_scheme* scheme, void* codes_host, void* scale_host, void* zp_host,
                                 float* global_scale_host);
int b200q_pipeline_sync(b200q_pipeline* p);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H */
