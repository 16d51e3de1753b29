"""AWQ scale search on the GPU: drop-in for llmcompressor ``AWQModifier._compute_best_scale`` and its helpers.

Reference behaviour (llmcompressor >= 0.9 modifiers/awq/base.py, restated in SURVEY.md Appendix A and
oracle/llmc_restated.py; driven from /root/reference/scripts/do_oneshot.py:179 with
configs/recipes/recipe_awq_w4a16.yaml):

    for ratio in i / n_grid:  s = x_mean^r / (w_mean^(1-r) + 1e-4) ... normalised;  for each balance layer:
        W <- fake_quantize(W * s) / s;   out = parent(x);   loss = sum((ref - out)^2) / numel;   keep first minimum

What changes here: |x| sums, w_mean, the per-ratio scale vectors, the scale->observe->fake-quantize->unscale weight
update and the squared-error reduction are single-pass CUDA kernels (csrc/awq_stats.cu, csrc/quant_group*.cu); losses
stay on the device (no per-sample ``.item()`` sync); with token-sharded calibration the |x| sums and the
``[n_grid]`` loss accumulators are all-reduced once each.  For a single-Linear parent the whole loss evaluation is
the fused tcgen05 kernel ``b200q_awq_gemm_loss`` (csrc/awq_gemm.cu).
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops


@torch.no_grad()
def abs_sum_cols(x: torch.Tensor, acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """_accumulate_mean numerator: acc[k] += sum_t |x[t, k]| (fp32 [K]); empty inputs are skipped."""
    L.require_cuda(x, acc)
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    if acc is None:
        acc = torch.zeros(x2.shape[1], dtype=torch.float32, device=x.device)
    if x2.shape[0]:
        L.check(L.lib().b200q_abs_sum_cols(L.ptr(x2), x2.shape[0], x2.shape[1], L.DTYPE_CODE[x2.dtype], L.ptr(acc), L.stream_ptr(x.device)))
    return acc


@torch.no_grad()
def compute_layer_means(weights: Sequence[torch.Tensor], group_size: int) -> torch.Tensor:
    """_compute_layer_means: mean over all balance-layer rows of |w| / (group_absmax + 1e-6) -> fp32 [K]."""
    K = weights[0].shape[1]
    acc = torch.zeros(K, dtype=torch.float64, device=weights[0].device)
    n = 0
    for w in weights:
        L.require_cuda(w)
        w = w.contiguous()
        L.check(L.lib().b200q_wmean_accumulate(L.ptr(w), w.shape[0], K, L.DTYPE_CODE[w.dtype], group_size, L.ptr(acc), L.stream_ptr(w.device)))
        n += w.shape[0]
    return (acc / n).float()


@torch.no_grad()
def awq_scales(x_mean: torch.Tensor, w_mean: Optional[torch.Tensor], ratios: Sequence[float], duo_scaling: bool) -> torch.Tensor:
    """Scale vectors of all grid points at once -> fp32 [n_ratios, K]."""
    L.require_cuda(x_mean, w_mean)
    K = x_mean.numel()
    out = torch.empty((len(ratios), K), dtype=torch.float32, device=x_mean.device)
    r = (ctypes.c_float * len(ratios))(*[float(v) for v in ratios])
    xm = x_mean.float().contiguous()
    wm = w_mean.float().contiguous() if w_mean is not None else None
    L.check(L.lib().b200q_awq_scales(L.ptr(xm), L.ptr(wm), K, r, len(ratios), int(bool(duo_scaling)), L.ptr(out), L.stream_ptr(x_mean.device)))
    return out


@torch.no_grad()
def scaled_fake_quantize(w: torch.Tensor, scales: torch.Tensor, args, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """W' = fake_quantize(W * s[None, :]) / s[None, :] with a fresh memoryless_minmax observer, one pass."""
    L.require_cuda(w, scales)
    w = w.contiguous()
    if out is None:
        out = torch.empty_like(w)
    sc = ops.scheme_from_args(args, w.dtype, True)
    L.check(L.lib().b200q_awq_scaled_fake_quantize(L.ptr(w), w.shape[0], w.shape[1], ctypes.byref(sc), L.ptr(scales.contiguous()),
                                                   L.ptr(out), L.stream_ptr(w.device)))
    return out


@torch.no_grad()
def scaled_fake_quantize_grid(w: torch.Tensor, scales: torch.Tensor, args, out: torch.Tensor) -> torch.Tensor:
    """All grid points at once: ``out[r] = fake_quantize(W * scales[r]) / scales[r]``; ``out`` is ``[R, rows, K]`` or a row slice
    ``buf[:, off:off + rows]`` of a larger ``[R, total_rows, K]`` buffer (the variants of several balance layers side by side)."""
    L.require_cuda(w, scales, out)
    w, scales = w.contiguous(), scales.contiguous()
    R, rows, K = out.shape
    if tuple(w.shape) != (rows, K) or tuple(scales.shape) != (R, K) or out.stride(2) != 1 or out.stride(1) != K:
        raise L.B200QError("scaled_fake_quantize_grid: operand shapes / strides do not agree")
    sc = ops.scheme_from_args(args, w.dtype, True)
    L.check(L.lib().b200q_awq_scaled_fake_quantize_grid(L.ptr(w), rows, K, ctypes.byref(sc), L.ptr(scales), R, L.ptr(out),
                                                        out.stride(0) if R > 1 else rows * K, L.stream_ptr(w.device)))
    return out


@torch.no_grad()
def sq_err_accumulate(y_ref: torch.Tensor, y_q: torch.Tensor, acc: torch.Tensor) -> None:
    """_compute_loss partial: acc[0] += sum((y_ref - y_q)^2), difference rounded to the output dtype first."""
    L.check(L.lib().b200q_sq_err_accumulate(L.ptr(y_ref.contiguous()), L.ptr(y_q.contiguous()), y_ref.numel(), L.DTYPE_CODE[y_ref.dtype],
                                            L.ptr(acc), L.stream_ptr(y_ref.device)))


class Workspace:
    """Persistent HBM scratch for the search (weight variants, projected activations): one flat buffer per role that only
    grows, so a layer-by-layer run re-uses the same 10-30 GB instead of cycling it through the caching allocator.
    One process drives one GPU and the search is sequential, so a single pool per device is enough."""

    def __init__(self):
        self._buf = {}

    def get(self, role: str, shape, dtype, device) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        key = (role, dtype, str(device))
        buf = self._buf.get(key)
        if buf is None or buf.numel() < n:
            self._buf[key] = buf = torch.empty(n, dtype=dtype, device=device)
        return buf[:n].view(*shape)

    def release(self):
        self._buf.clear()


workspace = Workspace()


def _ws(lib, T, K, N, R, dev):
    ws_bytes = int(lib.b200q_awq_gemm_loss_workspace(T, K, N, R))
    return torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev), ws_bytes


@torch.no_grad()
def gemm_loss_pairs(a_ref: torch.Tensor, a_q: Optional[torch.Tensor], b_ref: torch.Tensor, b_q: Optional[torch.Tensor]) -> torch.Tensor:
    """loss[r] = sum_{t,n} (bf16(a_ref b_ref^T) - bf16(A_r B_r^T))^2 with A_r = a_q[r] (or a_ref), B_r = b_q[r] (or b_ref), in one
    tcgen05 kernel; outputs are never materialised.  bf16 only.  Returns fp32 [R] (sums, not means)."""
    L.require_cuda(a_ref, a_q, b_ref, b_q)
    for t in (a_ref, a_q, b_ref, b_q):
        if t is not None and (t.dtype != torch.bfloat16 or not t.is_contiguous()):
            raise L.B200QError("gemm_loss_pairs takes contiguous bf16 tensors")
    if a_q is None and b_q is None:
        raise L.B200QError("gemm_loss_pairs: neither operand varies")
    T, K = a_ref.shape
    N = b_ref.shape[0]
    R = (a_q if a_q is not None else b_q).shape[0]
    if (a_q is not None and tuple(a_q.shape) != (R, T, K)) or (b_q is not None and tuple(b_q.shape) != (R, N, K)) or b_ref.shape[1] != K:
        raise L.B200QError("gemm_loss_pairs: operand shapes do not agree")
    lib = L.lib()
    ws, ws_bytes = _ws(lib, T, K, N, R, a_ref.device)
    loss = torch.zeros(R, dtype=torch.float32, device=a_ref.device)
    L.check(lib.b200q_awq_gemm_loss_pairs(L.ptr(a_ref), L.ptr(a_q), T, K, L.ptr(b_ref), L.ptr(b_q), N, R, L.ptr(loss), L.ptr(ws), ws_bytes,
                                          L.stream_ptr(a_ref.device)))
    return loss


@torch.no_grad()
def gemm_loss_fused(x: torch.Tensor, w_ref: torch.Tensor, w_q: torch.Tensor) -> torch.Tensor:
    """Single-Linear parent: loss[r] = sum_{t,n} (bf16(x w_ref^T) - bf16(x w_q[r]^T))^2 for all stacked variants ``w_q [R, N, K]``."""
    return gemm_loss_pairs(x.contiguous(), None, w_ref.contiguous(), w_q.contiguous())


@torch.no_grad()
def gemm_project(x: torch.Tensor, w_all: torch.Tensor, swiglu: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """First stage of a multi-layer parent for all weight variants ``w_all [V, rows, K]`` at once (tcgen05):
    ``out[v] = x w_all[v]^T`` (bf16), or with ``swiglu`` ``silu(x Wg[v]^T) * (x Wu[v]^T)`` where ``w_all[v] = [Wg; Wu]``."""
    L.require_cuda(x, w_all, out)
    if x.dtype != torch.bfloat16 or w_all.dtype != torch.bfloat16:
        raise L.B200QError("gemm_project takes bf16 tensors")
    x, w_all = x.contiguous(), w_all.contiguous()
    T, K = x.shape
    V, rows, _ = w_all.shape
    n_out = rows // 2 if swiglu else rows
    if out is None:
        out = torch.empty((V, T, n_out), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().b200q_awq_gemm_project(L.ptr(x), T, K, L.ptr(w_all), V, n_out, int(bool(swiglu)), L.ptr(out), L.stream_ptr(x.device)))
    return out


# ----------------------------------------------------------------------------- fused parents (W2 + W3 on the tensor cores)
class LinearParent:
    """Parent == the single balance Linear (up_proj -> down_proj, v_proj -> o_proj, per-expert w3 -> w2)."""

    def fused_losses(self, x: torch.Tensor, w_all: torch.Tensor) -> Tuple[torch.Tensor, int]:
        return gemm_loss_pairs(x, None, w_all[0], w_all[1:]), x.shape[0] * w_all.shape[1]

    def __call__(self, weights, x):
        return torch.nn.functional.linear(x, weights[0])


class MLPParent:
    """Parent == the gated MLP whose gate/up projections are the balance layers (post_attention_layernorm -> gate, up):
    out = down(silu(gate(x)) * up(x)).  Balance weights are stacked [gate; up]."""

    def __init__(self, down: torch.Tensor):
        self.down = down.contiguous()

    def fused_losses(self, x, w_all):
        h = gemm_project(x, w_all, swiglu=True, out=workspace.get("proj", (w_all.shape[0], x.shape[0], w_all.shape[1] // 2), x.dtype, x.device))
        return gemm_loss_pairs(h[0], h[1:], self.down, None), x.shape[0] * self.down.shape[0]

    def __call__(self, weights, x):
        F = torch.nn.functional
        return F.linear(F.silu(F.linear(x, weights[0])) * F.linear(x, weights[1]), self.down)


_ATTN_SDPA = __import__("os").environ.get("B200Q_ATTN_SDPA") is not None   # A/B switch: torch SDPA (library flash attention) instead of the tcgen05 core


def attention_core(qkv: torch.Tensor, n_heads: int, n_kv: int, head_dim: int, seq_len: int, q_norm, k_norm, cos, sin, eps=1e-6,
                   out: Optional[torch.Tensor] = None):
    """Qwen3 attention between the q/k/v projections and o_proj (transformers Qwen3Attention.forward): per-head RMSNorm
    on q and k, rotary embedding, causal GQA attention.  ``qkv [T, (H + 2 Hkv) d]`` with T = samples * seq_len is
    normalised / rotated IN PLACE by one CUDA kernel (``b200q_qk_norm_rope``); softmax(QK^T / sqrt(d)) V then runs per sample on
    the hand-written tcgen05 kernel ``b200q_attention_core`` (round 1 used torch SDPA here; ``B200Q_ATTN_SDPA=1`` still selects it
    for A/B runs).  Returns ``[T, H d]`` (written into ``out`` when given)."""
    L.require_cuda(qkv)
    T = qkv.shape[0]
    B = T // seq_len
    if qkv.dtype != torch.bfloat16 or not qkv.is_contiguous() or B * seq_len != T:
        raise L.B200QError("attention_core takes a contiguous bf16 [samples * seq_len, (H + 2 Hkv) d] tensor")
    L.check(L.lib().b200q_qk_norm_rope(L.ptr(qkv), T, n_heads, n_kv, head_dim, seq_len, L.ptr(q_norm), L.ptr(k_norm), L.ptr(cos), L.ptr(sin),
                                       ctypes.c_float(eps), L.stream_ptr(qkv.device)))
    if not _ATTN_SDPA:
        # softmax(Q K^T / sqrt(d), causal) V per sample on the tcgen05 kernel (csrc/awq_attn_core.cu)
        lib = L.lib()
        out = torch.empty((T, n_heads * head_dim), dtype=qkv.dtype, device=qkv.device) if out is None else out
        nws = int(lib.b200q_attention_workspace(T, n_kv, head_dim, seq_len))
        ws = workspace.get("attn_vt", (max(nws, 16),), torch.uint8, qkv.device)
        L.check(lib.b200q_attention_core(L.ptr(qkv), T, n_heads, n_kv, head_dim, seq_len, L.ptr(out), L.ptr(ws), nws, L.stream_ptr(qkv.device)))
        return out
    q, k, v = qkv.split([n_heads * head_dim, n_kv * head_dim, n_kv * head_dim], dim=-1)
    q = q.unflatten(-1, (n_heads, head_dim)).unflatten(0, (B, seq_len)).transpose(1, 2)
    k = k.unflatten(-1, (n_kv, head_dim)).unflatten(0, (B, seq_len)).transpose(1, 2)
    v = v.unflatten(-1, (n_kv, head_dim)).unflatten(0, (B, seq_len)).transpose(1, 2)
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=n_heads != n_kv)
    o = o.transpose(1, 2).reshape(T, n_heads * head_dim)
    if out is not None:
        out.copy_(o)
        return out
    return o


class AttentionParent:
    """Parent == self_attn whose q/k/v projections are the balance layers (input_layernorm -> q, k, v).  Balance weights
    are stacked [Wq; Wk; Wv].  Projections, the causal GQA attention core and the o_proj + loss all run on hand-written tcgen05
    kernels (awq_gemm.cu, awq_attn_core.cu)."""

    def __init__(self, o_proj: torch.Tensor, n_heads: int, n_kv: int, head_dim: int, seq_len: int, q_norm: torch.Tensor,
                 k_norm: torch.Tensor, rope_theta: float = 1e6, eps: float = 1e-6):
        self.o = o_proj.contiguous()
        self.cfg = (n_heads, n_kv, head_dim, seq_len)
        self.q_norm, self.k_norm, self.eps = q_norm.contiguous(), k_norm.contiguous(), eps
        inv = 1.0 / (rope_theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=o_proj.device) / head_dim))
        fr = torch.outer(torch.arange(seq_len, dtype=torch.float32, device=o_proj.device), inv)
        emb = torch.cat((fr, fr), dim=-1)
        self.cos, self.sin = emb.cos().to(o_proj.dtype).contiguous(), emb.sin().to(o_proj.dtype).contiguous()

    def core(self, qkv):
        return attention_core(qkv, *self.cfg, self.q_norm, self.k_norm, self.cos, self.sin, self.eps)

    def fused_losses(self, x, w_all):
        qkv = gemm_project(x, w_all, swiglu=False, out=workspace.get("proj", (w_all.shape[0], x.shape[0], w_all.shape[1]), x.dtype, x.device))
        attn = workspace.get("attn", (qkv.shape[0], x.shape[0], self.o.shape[1]), x.dtype, x.device)
        for v in range(qkv.shape[0]):
            attention_core(qkv[v], *self.cfg, self.q_norm, self.k_norm, self.cos, self.sin, self.eps, out=attn[v])
        return gemm_loss_pairs(attn[0], attn[1:], self.o, None), x.shape[0] * self.o.shape[0]

    def __call__(self, weights, x):
        F = torch.nn.functional
        qkv = torch.cat([F.linear(x, w) for w in weights], dim=-1).contiguous()
        return F.linear(self.core(qkv), self.o)


linear_parent = LinearParent()


def mlp_parent(down: torch.Tensor) -> MLPParent:
    return MLPParent(down)


def reduce_token_stats(xsum: torch.Tensor, count: torch.Tensor, group=None) -> torch.Tensor:
    """x_mean of token-sharded calibration: all-reduce(SUM) the per-channel |x| sums and the token count, then divide
    (``_accumulate_mean`` keeps exactly this running sum / count pair).  Works on any device / backend (NCCL, gloo)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(xsum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    return xsum / count.to(xsum.dtype)


def reduce_and_select(acc: torch.Tensor, group=None, distributed: bool = True) -> Tuple[int, List[float]]:
    """Finish ``_compute_best_scale``: ``acc = [sum_sq_err(ratio_0..n-1), numel]`` (fp32, this rank's token shard) is
    all-reduced (SUM) so every rank sees the same totals, the losses are sums / numel, and the FIRST minimum wins
    (upstream scans ``if loss < best_error``).  Raises when no ratio has a finite loss."""
    import torch.distributed as dist

    if distributed and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return _select_host(acc.double().cpu())


def _select_host(host: torch.Tensor) -> Tuple[int, List[float]]:
    n_grid = host.numel() - 1
    losses = (host[:n_grid] / host[n_grid]).tolist()
    best_err, best_i = float("inf"), -1
    for i, v in enumerate(losses):
        if v < best_err:
            best_err, best_i = v, i
    if best_i < 0:
        raise RuntimeError("AWQ: no finite loss for any ratio")
    return best_i, losses


def _first_min_device(acc: torch.Tensor) -> torch.Tensor:
    """Index of the first minimum of ``acc[:-1]`` (the scan ``if loss < best_error``; NaN never wins) as a device int64 [1], no
    host round trip; -1 when no loss is finite-or-smaller-than-inf (caught when the results are fetched)."""
    n = acc.numel() - 1
    v = torch.nan_to_num(acc[:n].double(), nan=float("inf"))
    idx = torch.arange(n, device=acc.device)
    m = v.min()
    first = torch.where(v == m, idx, torch.full_like(idx, n)).min()
    return torch.where(torch.isinf(m) & (m > 0), torch.full_like(first, -1), first).reshape(1)


@torch.no_grad()
def abs_mean_running(x: torch.Tensor, sample_len: int) -> torch.Tensor:
    """``x_mean_dtype="act"``: llmcompressor's ``_accumulate_mean`` as recalled from upstream (SURVEY.md Appendix A line 489) -- the
    hook keeps ``|x|`` in the ACTIVATION dtype, so each calibration batch contributes ``act(sum_t |x|)`` and the running mean
    ``(prev_mean * prev_count + batch_sum) / (prev_count + T_b)`` is evaluated in that dtype, batch after batch.  The per-batch
    sums come from ``b200q_abs_sum_cols`` (fp32 accumulation, rounded once like torch's reduction); the [K]-sized running update
    is a handful of elementwise ops per batch.  Order dependent, hence not token-shardable."""
    x2 = x.reshape(-1, x.shape[-1])
    mean, count = None, 0
    for t0 in range(0, x2.shape[0], sample_len):
        xb = x2[t0:t0 + sample_len]
        s = abs_sum_cols(xb).to(x.dtype)
        mean = s / xb.shape[0] if mean is None else (mean * count + s) / (count + xb.shape[0])
        count += xb.shape[0]
    return mean.float()


@torch.no_grad()
def _search_device(x: torch.Tensor, weights: Sequence[torch.Tensor], parent: Callable, args, n_grid: int, duo_scaling: bool,
                   process_group, fused: Optional[bool], token_chunk: int, x_mean_dtype: str = "fp32", loss_form: str = "float_pow",
                   sample_len: Optional[int] = None):
    """Everything of ``_compute_best_scale`` that runs on the device, enqueued without a host synchronisation:
    returns (scales fp32 [n_grid, K], acc fp32 [n_grid + 1] = summed squared errors + element count, ratios).

    Fidelity switches for the two places where llmcompressor versions are known to differ (defaults = the fused path):
      x_mean_dtype  "fp32" (|x| summed in fp32 over all tokens) | "act" (upstream's per-batch running mean in the activation dtype,
                    needs ``sample_len``; not token-shardable)
      loss_form     "float_pow" (``(a - b).float().pow(2).sum()``: bf16 difference, fp32 square / sum -- what the fused tcgen05
                    epilogue computes) | "mse_bf16" (``F.mse_loss(a, b, reduction="sum")`` per batch on the bf16 tensors: square and
                    per-batch result rounded to bf16; evaluated through the generic parent path, per ``sample_len`` tokens)"""
    import torch.distributed as dist

    if x_mean_dtype not in ("fp32", "act") or loss_form not in ("float_pow", "mse_bf16"):
        raise ValueError(f"unknown fidelity switch: x_mean_dtype={x_mean_dtype!r}, loss_form={loss_form!r}")
    if (x_mean_dtype == "act" or loss_form == "mse_bf16") and not sample_len:
        raise ValueError("x_mean_dtype='act' / loss_form='mse_bf16' follow the per-batch arithmetic and need sample_len")

    dev = x.device
    K = x.shape[-1]
    x = x.reshape(-1, K)
    # token-sharded calibration is explicit: layer- / expert-sharded ranks search different mappings and must NOT be reduced
    # together, so nothing is exchanged unless the caller names the group whose ranks hold shards of the same tokens
    dist_on = process_group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
    if loss_form == "mse_bf16":
        fused, token_chunk = False, int(sample_len)
    if fused is None:
        fused = hasattr(parent, "fused_losses") and x.dtype == torch.bfloat16
    # ---- statistics
    if x_mean_dtype == "act":
        if dist_on:
            raise ValueError("x_mean_dtype='act' is a sequential running mean: run it unsharded")
        x_mean = abs_mean_running(x, int(sample_len))
    else:
        xsum = abs_sum_cols(x)
        if dist_on:
            cnt = torch.full((1,), float(x.shape[0]), dtype=torch.float64, device=dev)
            x_mean = reduce_token_stats(xsum, cnt, process_group)
        else:
            x_mean = xsum / float(x.shape[0])
    w_mean = compute_layer_means(weights, args.group_size) if duo_scaling else None
    ratios = [i / n_grid for i in range(n_grid)]
    scales = awq_scales(x_mean, w_mean, ratios, duo_scaling)
    # ---- losses
    acc = torch.zeros(n_grid + 1, dtype=torch.float32, device=dev)  # [n_grid] sums + numel
    if fused:
        if not hasattr(parent, "fused_losses"):
            raise L.B200QError("fused evaluation needs a LinearParent / MLPParent / AttentionParent")
        rows = [w.shape[0] for w in weights]
        w_all = workspace.get("w_all", (n_grid + 1, sum(rows), K), weights[0].dtype, dev)  # [0] = reference weights
        off = 0
        for w, n in zip(weights, rows):
            w_all[0, off:off + n].copy_(w)
            scaled_fake_quantize_grid(w, scales, args, w_all[1:, off:off + n])  # all ratios in one launch
            off += n
        sums, numel = parent.fused_losses(x.contiguous(), w_all)
        acc[:n_grid] = sums
        acc[n_grid:].fill_(float(numel))  # fill_ takes the scalar by value: `acc[i] = python_float` stages a host copy and blocks
    else:
        wq = [torch.empty_like(w) for w in weights]
        chunks = [x[t0:t0 + token_chunk] for t0 in range(0, x.shape[0], token_chunk)]
        refs = [parent(list(weights), xc) for xc in chunks]
        numel = sum(r.numel() for r in refs)
        for i in range(n_grid):
            for w, o in zip(weights, wq):
                scaled_fake_quantize(w, scales[i], args, out=o)
            for xc, ref in zip(chunks, refs):
                if loss_form == "mse_bf16":
                    acc[i:i + 1] += torch.nn.functional.mse_loss(ref, parent(wq, xc), reduction="sum").float()
                else:
                    sq_err_accumulate(ref, parent(wq, xc), acc[i:i + 1])
        acc[n_grid:].fill_(float(numel))  # fill_ takes the scalar by value: `acc[i] = python_float` stages a host copy and blocks
    if dist_on:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=process_group)
    return scales, acc, ratios


@torch.no_grad()
def compute_best_scale(x: torch.Tensor, weights: Sequence[torch.Tensor], parent: Callable, args, n_grid: int = 20,
                       duo_scaling: bool = True, process_group=None, fused: Optional[bool] = None, fused_linear: Optional[bool] = None,
                       token_chunk: int = 8192, x_mean_dtype: str = "fp32", loss_form: str = "float_pow",
                       sample_len: Optional[int] = None) -> Tuple[torch.Tensor, float, List[float]]:
    """``AWQModifier._compute_best_scale`` for one mapping.

    x        [T_local, K] inputs of the balance layers (this rank's token shard; whole samples for an attention parent)
    weights  balance-layer weights [N_i, K];  parent(weights, x_chunk) -> parent-module output
    parent   a LinearParent / MLPParent / AttentionParent (fused tensor-core evaluation, bf16) or any callable (generic
             evaluation: the parent runs through torch, the squared error through ``b200q_sq_err_accumulate``)
    Returns (best_scales fp32 [K] on the CPU like the reference, best_ratio, losses[n_grid]).  Raises if no ratio gives
    a finite loss.  With ``process_group`` (token-sharded calibration; pass ``dist.group.WORLD`` for all ranks) the |x| sums /
    token counts and the loss accumulators are all-reduced (SUM), so every rank of the group returns the same argmin."""
    if fused is None and fused_linear is not None:
        fused = fused_linear
    scales, acc, ratios = _search_device(x, weights, parent, args, n_grid, duo_scaling, process_group, fused, token_chunk,
                                         x_mean_dtype, loss_form, sample_len)
    best_i, losses = _select_host(acc.double().cpu())
    return scales[best_i].cpu(), ratios[best_i], losses


@torch.no_grad()
def smooth(weights: Sequence[torch.Tensor], smooth_weight: torch.Tensor, scales: torch.Tensor):
    """_smooth: balance W *= s[None, :] in place; smooth layer /= s (1-D norm weight or last len(s) rows)."""
    s = scales.to(weights[0].device, torch.float32)
    for w in weights:
        w.copy_((w.float() * s.view(1, -1)).to(w.dtype))
    if smooth_weight.ndim == 1:
        smooth_weight.copy_((smooth_weight.float() / s.to(smooth_weight.device)).to(smooth_weight.dtype))
    else:
        k = s.numel()
        smooth_weight[-k:].copy_((smooth_weight[-k:].float() / s.to(smooth_weight.device).view(-1, 1)).to(smooth_weight.dtype))


# ----------------------------------------------------------------------------- one dense decoder layer (config 1)
def decoder_layer_flops(T: int, hidden: int, inter: int, n_heads: int, n_kv: int, head_dim: int, seq_len: int, n_grid: int = 20) -> float:
    """ALGORITHMIC FLOPs of the AWQ search of one dense decoder layer: (n_grid + 1) parent evaluations of
    the q/k/v mapping (q,k,v,o projections + dense attention), the gate/up mapping (gate, up, down) and the down mapping."""
    qkv = 2.0 * T * hidden * (n_heads + 2 * n_kv) * head_dim
    o = 2.0 * T * n_heads * head_dim * hidden
    attn = 4.0 * seq_len * seq_len * head_dim * n_heads * (T // seq_len)
    mlp = 3 * 2.0 * T * hidden * inter
    down = 2.0 * T * hidden * inter
    return (n_grid + 1) * (qkv + o + attn + mlp + down)


@torch.no_grad()
def search_decoder_layer(weights: dict, acts: dict, args, n_heads: int, n_kv: int, head_dim: int, seq_len: int, n_grid: int = 20,
                         duo_scaling: bool = True, apply: bool = True, process_group=None, x_mean_dtype: str = "fp32",
                         loss_form: str = "float_pow") -> dict:
    """AWQ scale search of one dense decoder layer with llmcompressor's default Llama/Qwen mappings
    (LLMC modifiers/awq/mappings.py; REF:configs/recipes/recipe_awq_w4a16.yaml uses the defaults):

        input_layernorm          -> q_proj, k_proj, v_proj     parent self_attn
        v_proj                   -> o_proj                     skipped under GQA (v rows != o cols), as upstream does
        post_attention_layernorm -> gate_proj, up_proj         parent mlp
        up_proj                  -> down_proj                  parent down_proj

    weights: q,k,v,o,gate,up,down [N,K] + input_layernorm, post_attention_layernorm, q_norm, k_norm (1-D)
    acts:    "attn_in" [T, hidden], "mlp_in" [T, hidden], "down_in" [T, inter]   (inputs of the balance layers)
    With ``apply`` the best scales are folded in like ``_smooth`` (balance W *= s, smooth layer /= s) before the next mapping.
    Returns {mapping: (best_scales cpu, best_ratio, losses)}."""
    out = {}
    w = weights
    fid = dict(x_mean_dtype=x_mean_dtype, loss_form=loss_form, sample_len=seq_len, token_chunk=seq_len * max(1, 8192 // seq_len))
    attn = AttentionParent(w["o"], n_heads, n_kv, head_dim, seq_len, w["q_norm"], w["k_norm"])
    out["qkv"] = compute_best_scale(acts["attn_in"], [w["q"], w["k"], w["v"]], attn, args, n_grid, duo_scaling, process_group, **fid)
    if apply:
        smooth([w["q"], w["k"], w["v"]], w["input_layernorm"], out["qkv"][0])
    if w["v"].shape[0] == w["o"].shape[1]:
        out["v_o"] = compute_best_scale(acts["o_in"], [w["o"]], linear_parent, args, n_grid, duo_scaling, process_group, **fid)
        if apply:
            smooth([w["o"]], w["v"], out["v_o"][0])
    out["gate_up"] = compute_best_scale(acts["mlp_in"], [w["gate"], w["up"]], MLPParent(w["down"]), args, n_grid, duo_scaling, process_group, **fid)
    if apply:
        smooth([w["gate"], w["up"]], w["post_attention_layernorm"], out["gate_up"][0])
    out["down"] = compute_best_scale(acts["down_in"], [w["down"]], linear_parent, args, n_grid, duo_scaling, process_group, **fid)
    if apply:
        smooth([w["down"]], w["up"], out["down"][0])
    return out


# ----------------------------------------------------------------------------- MoE experts (config 5)
def expert_mapping_flops(T: int, k: int, n: int, n_grid: int = 20) -> float:
    """ALGORITHMIC FLOPs of one per-expert single-Linear mapping search: (n_grid + 1) evaluations of x [T,k] @ W[n,k]^T."""
    return (n_grid + 1) * 2.0 * T * k * n


@torch.no_grad()
def search_expert_mappings(x_in: Sequence[torch.Tensor], balance: torch.Tensor, args, n_grid: int = 20, duo_scaling: bool = True,
                           smooth_weight: Optional[torch.Tensor] = None, process_group=None) -> List[Tuple[torch.Tensor, float, List[float]]]:
    """The per-expert mappings of a MoE layer whose parent is the balance Linear itself -- ``w3 -> w2`` in
    REF:configs/recipes/recipe_Minimax-M2.1-Experts-only-AWQ.yaml:29-34 (``up_proj -> down_proj`` for Qwen3-MoE): one
    independent ``_compute_best_scale`` per expert, the embarrassingly expert-parallel part of SURVEY.md §8d config 5 (ii).

    x_in          per local expert the input of its balance layer ``[T, K]`` (= act(w1 x) * (w3 x); under
                  ``moe_calibrate_all_experts`` every expert sees all T tokens, REF:scripts/do_oneshot.py:186)
    balance       the local experts' balance weights stacked ``[E_local, N, K]`` (w2)
    smooth_weight optional ``[E_local, K, H]`` smooth layers (w3): when given the best scales are applied like ``_smooth``
                  (balance ``*= s``, the smooth layer's rows ``/= s``)
    Experts are sharded across ranks by the caller (``scheduler.partition``); nothing is exchanged between expert shards.
    ``process_group`` names ranks holding token shards of the SAME experts (statistics and losses are all-reduced over it).
    Returns one (best_scales cpu fp32 [K], best_ratio, losses[n_grid]) per local expert."""
    if len(x_in) != balance.shape[0]:
        raise L.B200QError("search_expert_mappings: one input per stacked expert weight is needed")
    if len(x_in) == 0:      # a rank that owns no expert of this layer (more ranks than experts)
        return []
    # everything stays on the device until all experts are enqueued: no host round trip (and no idle GPU) between experts
    pend = []
    for e, x in enumerate(x_in):
        scales, acc, ratios = _search_device(x, [balance[e]], linear_parent, args, n_grid, duo_scaling, process_group, None, 8192)
        best = _first_min_device(acc)
        best_scales = scales.index_select(0, best.clamp(min=0)).reshape(-1)
        if smooth_weight is not None:
            smooth([balance[e]], smooth_weight[e], best_scales)
        pend.append((best_scales, best, acc, ratios))
    out = []
    all_best = torch.cat([p[1] for p in pend]).cpu()
    all_scales = torch.stack([p[0] for p in pend]).cpu()
    all_acc = torch.stack([p[2] for p in pend]).double().cpu()
    for e, (_, _, _, ratios) in enumerate(pend):
        bi = int(all_best[e])
        if bi < 0:
            raise RuntimeError(f"AWQ: no finite loss for any ratio (expert {e})")
        n = all_acc.shape[1] - 1
        out.append((all_scales[e], ratios[bi], (all_acc[e, :n] / all_acc[e, n]).tolist()))
    return out


def moe_block_flops(T: int, top_k: int, hidden: int, inter: int, n_grid: int = 20) -> float:
    """ALGORITHMIC FLOPs of the layer-wide MoE mapping search: (n_grid + 1) evaluations of the routed block, 3 GEMMs of
    2 * hidden * inter per routed (token, expert) pair (un-routed pairs do not reach the block output)."""
    return (n_grid + 1) * 3 * 2.0 * T * top_k * hidden * inter


class RoutedMoE:
    """The routed sparse-MoE block as an AWQ parent: ``out[t] = sum_{j < k} p[t, j] * w2_e( act(w1_e x_t) * (w3_e x_t) )`` with
    ``e = topk_idx[t, j]``, accumulated expert by expert in the activation dtype like the transformers Mixtral / Qwen3-MoE /
    MiniMax blocks do (``index_add_`` per expert, ascending).  Routing depends on x and the (un-quantized) router only, so it is
    computed once by the caller.  Under ``moe_calibrate_all_experts`` every expert additionally runs on all tokens, but only
    routed pairs reach the block output that the loss is taken on, so only those are evaluated here (same output, 1 / (E / k) of
    the FLOPs).

    Layout: the routed (token, expert) pairs are sorted expert-major and every expert's rows are padded to whole 128-row tiles,
    so ONE grouped tcgen05 launch per stage serves all experts (``b200q_awq_gemm_project_grouped``: tile -> expert table); the
    combine kernel then sums each token's k rows in ascending expert order with bf16 rounding after every step."""

    TILE = 128

    def __init__(self, x: torch.Tensor, w2: torch.Tensor, topk_idx: torch.Tensor, topk_w: torch.Tensor, expert_offset: int = 0):
        """``expert_offset``: expert-parallel use -- ``w2`` holds experts ``[expert_offset, expert_offset + w2.shape[0])`` of the layer
        and only the pairs routed to them are laid out; a token's other slots get row -1 (skipped by the combine kernel)."""
        L.require_cuda(x, w2, topk_idx, topk_w)
        T, k = topk_idx.shape
        E = w2.shape[0]
        dev = x.device
        flat = topk_idx.reshape(-1).to(torch.int64) - int(expert_offset)
        local = (flat >= 0) & (flat < E)
        key = torch.where(local, flat, torch.full_like(flat, E))       # foreign pairs sort behind every local expert
        order = torch.argsort(key, stable=True)                       # pairs expert-major, token order inside an expert
        counts = torch.bincount(key, minlength=E + 1)[:E]
        padded = (counts + self.TILE - 1) // self.TILE * self.TILE
        starts = torch.cumsum(padded, 0) - padded                     # first padded row of each expert
        rows_total = int(padded.sum().item())                         # the one host round trip of the mapping
        rows_total = max(rows_total, self.TILE)
        excl = torch.cumsum(counts, 0) - counts
        sorted_e = key[order]
        sorted_local = sorted_e < E
        se = sorted_e.clamp(max=E - 1)
        prow = starts[se] + (torch.arange(order.numel(), device=dev) - excl[se])                 # padded row of each sorted pair
        prow = torch.where(sorted_local, prow, torch.full_like(prow, -1))
        gather = torch.zeros(rows_total + 1, dtype=torch.int64, device=dev)                      # padding rows read token 0
        gather[torch.where(sorted_local, prow, torch.full_like(prow, rows_total))] = order // k   # foreign pairs land in the spare slot
        self.xs = x.index_select(0, gather[:rows_total])               # [rows_total, H] routed inputs
        tile_e = torch.full((rows_total // self.TILE,), -1, dtype=torch.int32, device=dev)
        tile_ids = torch.arange(rows_total // self.TILE, device=dev) * self.TILE
        owner = torch.searchsorted(starts + padded, tile_ids, right=True).clamp(max=E - 1)
        valid = tile_ids < (starts + counts)[owner]                     # tiles that hold at least one real row
        tile_e[valid] = owner[valid].to(torch.int32)
        self.tile_expert = tile_e
        # per token: its k padded rows and routing weights in ascending expert order
        row_of_pair = torch.empty(order.numel(), dtype=torch.int64, device=dev)
        row_of_pair[order] = prow
        row_tk = row_of_pair.view(T, k)
        by_expert = torch.argsort(topk_idx.to(torch.int64), dim=1, stable=True)
        self.rows = row_tk.gather(1, by_expert).to(torch.int32).contiguous()
        self.pw = topk_w.to(x.dtype).gather(1, by_expert).contiguous()
        self.w2, self.T, self.H, self.k = w2.contiguous(), T, x.shape[1], k

    def project(self, w13: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The two grouped GEMM stages: w13 ``[E, 2 * I, H]`` -> per-pair expert outputs ``y [padded rows, H]``."""
        lib = L.lib()
        dev = self.xs.device
        st = L.stream_ptr(dev)
        E, two_i, H = w13.shape
        I = two_i // 2
        R = self.xs.shape[0]
        h = workspace.get("moe_h", (R, I), self.xs.dtype, dev)
        if y is None:
            y = workspace.get("moe_y", (R, self.H), self.xs.dtype, dev)
        L.check(lib.b200q_awq_gemm_project_grouped(L.ptr(self.xs), R, H, L.ptr(w13), E, I, 1, L.ptr(self.tile_expert), L.ptr(h), st))
        L.check(lib.b200q_awq_gemm_project_grouped(L.ptr(h), R, I, L.ptr(self.w2), E, self.H, 0, L.ptr(self.tile_expert), L.ptr(y), st))
        return y

    def combine(self, y: torch.Tensor, out: Optional[torch.Tensor] = None, init: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``out[t] = init[t] + sum_j bf16(y[rows[t, j]] * p[t, j])`` in ascending expert order, bf16 rounding after every step
        (``init`` None: zeros; it may be ``out`` itself)."""
        if out is None:
            out = torch.empty((self.T, self.H), dtype=self.xs.dtype, device=self.xs.device)
        L.check(L.lib().b200q_moe_combine_acc(L.ptr(y), L.ptr(self.rows), L.ptr(self.pw), self.T, self.k, self.H, L.ptr(init), L.ptr(out),
                                              L.stream_ptr(self.xs.device)))
        return out

    def __call__(self, w13: torch.Tensor) -> torch.Tensor:
        """w13 ``[E, 2 * I, H]`` (w1 rows, then w3 rows) -> block output ``[T, H]``."""
        return self.combine(self.project(w13))


@torch.no_grad()
def search_moe_block_mapping(x: torch.Tensor, w1: torch.Tensor, w3: torch.Tensor, w2: torch.Tensor, topk_idx: torch.Tensor,
                             topk_w: torch.Tensor, args, n_grid: int = 20, duo_scaling: bool = True, process_group=None,
                             max_variant_bytes: int = 16 << 30) -> Tuple[torch.Tensor, float, List[float]]:
    """The layer-wide MoE mapping ``post_attention_layernorm -> every expert's w1, w3`` (REF:configs/recipes/
    recipe_Minimax-M2.1-Experts-only-AWQ.yaml:29-31; Qwen3-MoE's ``post_attention_layernorm -> experts.*.gate/up``): ONE scale
    vector ``s[H]`` shared by all 2E balance matrices, ``w_mean`` over all their rows, ``x_mean`` over all tokens (the hook sits on
    expert 0's w1, which sees every token under calibrate-all-experts), parent = the routed block (``RoutedMoE``), loss on its
    output.

    x [T, H] bf16, w1 / w3 [E, I, H], w2 [E, H, I], topk_idx / topk_w [T, k].  The weight variants of a ratio are 2 * E * I * H
    elements, so ratios are processed in chunks of at most ``max_variant_bytes``.  With ``process_group`` the ranks hold token
    shards of the same layer (all experts on every rank): |x| sums and the loss accumulators are all-reduced.
    Returns (best_scales cpu fp32 [H], best_ratio, losses[n_grid]); the caller applies ``smooth`` to w1 / w3 / the norm."""
    import torch.distributed as dist

    L.require_cuda(x, w1, w3, w2)
    if x.dtype != torch.bfloat16:
        raise L.B200QError("search_moe_block_mapping runs on the bf16 tensor-core path")
    E, I, H = w1.shape
    dev = x.device
    dist_on = process_group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
    w13 = torch.cat([w1, w3], dim=1).contiguous()              # [E, 2I, H]
    xsum = abs_sum_cols(x)
    if dist_on:
        x_mean = reduce_token_stats(xsum, torch.full((1,), float(x.shape[0]), dtype=torch.float64, device=dev), process_group)
    else:
        x_mean = xsum / float(x.shape[0])
    rows_per_call = 262144  # launch limit of the row-tiled kernels
    flat = w13.view(-1, H)
    w_mean = compute_layer_means([flat[r0:r0 + rows_per_call] for r0 in range(0, flat.shape[0], rows_per_call)], args.group_size) if duo_scaling else None
    ratios = [i / n_grid for i in range(n_grid)]
    scales = awq_scales(x_mean, w_mean, ratios, duo_scaling)
    parent = RoutedMoE(x, w2, topk_idx, topk_w)
    ref = parent(w13)
    acc = torch.zeros(n_grid + 1, dtype=torch.float32, device=dev)
    per_ratio = w13.numel() * w13.element_size()
    chunk = max(1, min(n_grid, int(max_variant_bytes // max(per_ratio, 1))))
    for r0 in range(0, n_grid, chunk):
        r1 = min(n_grid, r0 + chunk)
        variants = workspace.get("moe_w13", (r1 - r0, E, 2 * I, H), w13.dtype, dev)
        vflat = variants.view(r1 - r0, E * 2 * I, H)
        for q0 in range(0, flat.shape[0], rows_per_call):
            scaled_fake_quantize_grid(flat[q0:q0 + rows_per_call], scales[r0:r1], args, vflat[:, q0:q0 + rows_per_call])
        for r in range(r0, r1):
            sq_err_accumulate(ref, parent(variants[r - r0]), acc[r:r + 1])
    acc[n_grid:].fill_(float(ref.numel()))
    best_i, losses = reduce_and_select(acc, process_group if dist_on else None, dist_on)
    return scales[best_i].cpu(), ratios[best_i], losses


def _all_gather_rows(t: torch.Tensor, group, sizes: Optional[List[int]] = None) -> torch.Tensor:
    """Concatenate the ranks' row shards (rank order; shards may differ in length).  ``sizes``: the ranks' row counts when they are
    already known (``_gather_sizes``), saving the exchange."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if sizes is None:
        sizes = _gather_sizes(t.shape[0], t.device, group)
    m = max(sizes)
    pad = t if t.shape[0] == m else torch.cat([t, t.new_zeros((m - t.shape[0],) + tuple(t.shape[1:]))])
    if min(sizes) == m:
        out = torch.empty((world * m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
        return out
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])


def _gather_sizes(n_rows: int, device, group) -> List[int]:
    import torch.distributed as dist

    n = torch.tensor([n_rows], dtype=torch.int64, device=device)
    sizes = torch.empty(dist.get_world_size(group), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    return [int(v) for v in sizes.tolist()]


@torch.no_grad()
def search_moe_block_mapping_ep(x: torch.Tensor, w1: torch.Tensor, w3: torch.Tensor, w2: torch.Tensor, topk_idx: torch.Tensor,
                                topk_w: torch.Tensor, args, expert_offset: int, process_group, n_grid: int = 20,
                                duo_scaling: bool = True, max_variant_bytes: int = 16 << 30,
                                ring_chunks: int = 4) -> Tuple[torch.Tensor, float, List[float]]:
    """``search_moe_block_mapping`` with the EXPERTS partitioned over the ranks (rank r holds the ascending range
    ``[expert_offset, expert_offset + w1.shape[0])``; ranges in rank order) and each rank's token shard ``x`` / ``topk_*`` all-gathered.

    Why not token sharding: every rank would fake-quantise all 2E matrices for all ratios (replicated work: 49 of 96 ms on 8 GPUs at
    256 experts).  Here a rank scales, fake-quantises and multiplies only its own experts, for all tokens.  The block output is a
    per-token sum over the token's experts accumulated IN ASCENDING EXPERT ORDER with bf16 rounding after every step (transformers'
    per-expert ``index_add_``), so the partial outputs are not summed by a reduction (a different rounding sequence: ~2 % loss noise)
    but passed along the ring 0 -> 1 -> ... -> N-1: rank r receives the running bf16 output of the experts before its own, continues
    the sequence with ``b200q_moe_combine_acc`` and sends it on; the last rank holds the output bit-identical to the single-GPU one
    and takes the loss.  One ring pass per evaluated weight set (reference + n_grid ratios), NCCL send/recv on a side stream, the
    GEMMs of later passes running ahead on the compute stream.  The [n_grid] losses are all-reduced (only the last rank's are non-zero).
    Returns (best_scales cpu fp32 [H], best_ratio, losses[n_grid]) on every rank."""
    import torch.distributed as dist

    L.require_cuda(x, w1, w3, w2)
    if x.dtype != torch.bfloat16:
        raise L.B200QError("search_moe_block_mapping_ep runs on the bf16 tensor-core path")
    if process_group is None or not dist.is_initialized() or dist.get_world_size(process_group) < 2:
        raise L.B200QError("search_moe_block_mapping_ep needs a process group of at least 2 ranks (one rank: search_moe_block_mapping)")
    world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    prev_rank = dist.get_global_rank(process_group, rank - 1) if rank > 0 else None
    next_rank = dist.get_global_rank(process_group, rank + 1) if rank + 1 < world else None
    # hop r -> r + 1 runs on ring group r % 2: a rank's receive (from r - 1) and its send (to r + 1) then sit on different NCCL
    # communicators / streams and overlap (unbatched P2P ops on ONE eagerly initialised group are serialised with each other)
    hop_groups = _ring_groups(process_group) if os.environ.get("B200Q_EP_RING", "symm") == "nccl" else (process_group, process_group)
    recv_group, send_group = hop_groups[(rank - 1) % 2], hop_groups[rank % 2]
    E, I, H = w1.shape
    dev = x.device
    # the ring order must be the expert order: rank r's range starts where rank r - 1's ends
    span = torch.tensor([int(expert_offset), int(expert_offset) + E], dtype=torch.int64, device=dev)
    spans = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(spans, span, group=process_group)
    spans = spans.view(world, 2).tolist()
    if any(b <= a for a, b in spans):
        raise L.B200QError(f"search_moe_block_mapping_ep: every rank needs at least one expert, got ranges {spans}")
    if spans[0][0] != 0 or any(spans[r][0] != spans[r - 1][1] for r in range(1, world)):
        raise L.B200QError(f"search_moe_block_mapping_ep: expert ranges must be contiguous and ascending in rank order, got {spans}")
    # ---- every rank sees all tokens (activations are small next to the weights: T x H bf16)
    sizes = _gather_sizes(x.shape[0], dev, process_group)
    x = _all_gather_rows(x.contiguous(), process_group, sizes)
    topk_idx = _all_gather_rows(topk_idx.contiguous(), process_group, sizes)
    topk_w = _all_gather_rows(topk_w.contiguous(), process_group, sizes)
    T = x.shape[0]
    w13 = torch.cat([w1, w3], dim=1).contiguous()              # [E, 2I, H], this rank's experts
    x_mean = abs_sum_cols(x) / float(T)                         # the same bits on every rank, and the single-GPU bits
    rows_per_call = 262144  # launch limit of the row-tiled kernels
    flat = w13.view(-1, H)
    w_mean = None
    if duo_scaling:
        acc64 = torch.zeros(H, dtype=torch.float64, device=dev)
        for r0 in range(0, flat.shape[0], rows_per_call):
            blk = flat[r0:r0 + rows_per_call]
            L.check(L.lib().b200q_wmean_accumulate(L.ptr(blk), blk.shape[0], H, L.DTYPE_CODE[blk.dtype], args.group_size, L.ptr(acc64),
                                                   L.stream_ptr(dev)))
        n_rows = torch.tensor([float(flat.shape[0])], dtype=torch.float64, device=dev)
        dist.all_reduce(acc64, op=dist.ReduceOp.SUM, group=process_group)
        dist.all_reduce(n_rows, op=dist.ReduceOp.SUM, group=process_group)
        w_mean = (acc64 / n_rows).float()
    ratios = [i / n_grid for i in range(n_grid)]
    scales = awq_scales(x_mean, w_mean, ratios, duo_scaling)
    parent = RoutedMoE(x, w2, topk_idx, topk_w, expert_offset=expert_offset)
    R = parent.xs.shape[0]
    n_pass = n_grid + 1                                         # pass 0: the unquantised weights, pass 1 + i: ratio i
    ys = workspace.get("moe_ep_y", (n_pass, R, H), x.dtype, dev)
    # The running outputs live in symmetric memory when the ranks can map each other's buffers (one node, NVLink): the combine kernel
    # then writes its result straight into the NEXT rank's buffer -- the transfer is the kernel's own stores over NVLink, overlapped
    # with the GEMMs of later passes, and a flag in the peer's signal pad hands the chunk over.  Measured at N = 8 (256 experts, 32 768
    # tokens): NCCL send/recv moved a 201 MB pass in ~2.2 ms per hop and bounded the search at 53 ms; B200Q_EP_RING=nccl keeps it.
    symm = _ep_symm_buffer(n_pass * T * H, x.dtype, dev, process_group) if os.environ.get("B200Q_EP_RING", "symm") != "nccl" else None
    if symm is not None:
        outs = symm[0].view(n_pass, T, H)
        peer_outs = symm[1].get_buffer(rank + 1, (n_pass, T, H), x.dtype) if rank + 1 < world else None
    else:
        outs = workspace.get("moe_ep_out", (n_pass, T, H), x.dtype, dev)
    acc = torch.zeros(n_grid + 1, dtype=torch.float32, device=dev)
    main = torch.cuda.current_stream(dev)
    side = _ep_side_stream(dev)
    side.wait_stream(main)
    y_ready = [torch.cuda.Event() for _ in range(n_pass)]
    sends = []

    # a pass travels in token chunks: the ring's fill / drain time is (N - 1) hops of ONE chunk instead of the whole [T, H] output
    n_chunks = max(1, min(ring_chunks, T // 1024))
    cuts = [T * c // n_chunks for c in range(n_chunks + 1)]

    def ring_step(p):
        # runs on the side stream: wait for this pass's GEMMs, take over the running output, add our experts, pass it on
        with torch.cuda.stream(side):
            side.wait_event(y_ready[p])
            for c in range(n_chunks):
                t0, t1 = cuts[c], cuts[c + 1]
                out = outs[p, t0:t1]
                channel = (p * n_chunks + c) % 8
                if prev_rank is not None:
                    if symm is not None:
                        symm[1].wait_signal(rank - 1, channel, 120000)
                    else:
                        dist.recv(out, src=prev_rank, group=recv_group)
                dst = peer_outs[p, t0:t1] if (symm is not None and next_rank is not None) else out
                L.check(L.lib().b200q_moe_combine_acc(L.ptr(ys[p]), L.ptr(parent.rows[t0:t1]), L.ptr(parent.pw[t0:t1]), t1 - t0, parent.k, H,
                                                      L.ptr(out) if prev_rank is not None else None, L.ptr(dst), L.stream_ptr(dev)))
                if next_rank is not None:
                    if symm is not None:
                        symm[1].put_signal(rank + 1, channel, 120000)
                    else:
                        sends.append(dist.isend(out, dst=next_rank, group=send_group))
            if next_rank is None and p > 0:
                sq_err_accumulate(outs[0], outs[p], acc[p - 1:p])

    parent.project(w13, ys[0])
    y_ready[0].record(main)
    ring_step(0)
    per_ratio = w13.numel() * w13.element_size()
    chunk = max(1, min(n_grid, int(max_variant_bytes // max(per_ratio, 1))))
    for r0 in range(0, n_grid, chunk):
        r1 = min(n_grid, r0 + chunk)
        variants = workspace.get("moe_w13", (r1 - r0, E, 2 * I, H), w13.dtype, dev)
        vflat = variants.view(r1 - r0, E * 2 * I, H)
        for q0 in range(0, flat.shape[0], rows_per_call):
            scaled_fake_quantize_grid(flat[q0:q0 + rows_per_call], scales[r0:r1], args, vflat[:, q0:q0 + rows_per_call])
        for r in range(r0, r1):
            parent.project(variants[r - r0], ys[1 + r])
            y_ready[1 + r].record(main)
            ring_step(1 + r)
    main.wait_stream(side)
    for w in sends:
        w.wait()
    if next_rank is None:
        acc[n_grid:].fill_(float(T * H))
    best_i, losses = reduce_and_select(acc, process_group, True)
    return scales[best_i].cpu(), ratios[best_i], losses


_EP_SIDE = {}
_RING_GROUPS = {}
_EP_SYMM = {}


def _ep_symm_buffer(numel: int, dtype, dev, process_group):
    """(tensor, handle) of a symmetric-memory buffer of ``numel`` elements shared by the ranks of ``process_group`` (collective; cached
    per group, re-made when the size changes), or None when the ranks cannot map each other's memory -- the caller then uses NCCL
    send/recv.  Every rank takes the same branch: the outcome is agreed with a MIN all-reduce."""
    import torch.distributed as dist

    key = id(process_group)
    hit = _EP_SYMM.get(key)
    if hit is not None and hit[0].numel() == numel and hit[0].dtype == dtype:
        return hit
    _EP_SYMM.pop(key, None)
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    t = hdl = None
    try:
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty((numel,), dtype=dtype, device=dev)
    except Exception:
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
    if int(ok.item()) == 0:
        return None
    try:
        hdl = symm_mem.rendezvous(t, process_group)
        if hdl.world_size != dist.get_world_size(process_group) or hdl.rank != dist.get_rank(process_group):
            raise RuntimeError("symmetric memory group mismatch")
    except Exception:
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
    if int(ok.item()) == 0:
        return None
    _EP_SYMM[key] = (t, hdl)
    return _EP_SYMM[key]


def _ring_groups(process_group):
    """Two extra process groups over the ranks of ``process_group`` (created once, collectively, in the same order on every rank)."""
    import torch.distributed as dist

    key = id(process_group)
    if key not in _RING_GROUPS:
        ranks = dist.get_process_group_ranks(process_group)
        _RING_GROUPS[key] = (dist.new_group(ranks=ranks), dist.new_group(ranks=ranks))
    return _RING_GROUPS[key]


def _ep_side_stream(dev) -> "torch.cuda.Stream":
    key = torch.device(dev).index
    if key not in _EP_SIDE:
        _EP_SIDE[key] = torch.cuda.Stream(dev)
    return _EP_SIDE[key]
