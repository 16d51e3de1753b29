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
def sq_err_accumulate(y_ref: torch.Tensor, y_q: torch.Tensor, acc: torch.Tensor) -> None:
    """_compute_loss partial: acc[0] += sum((y_ref - y_q)^2), difference rounded to the output dtype first."""
    L.check(L.lib().b200q_sq_err_accumulate(L.ptr(y_ref.contiguous()), L.ptr(y_q.contiguous()), y_ref.numel(), L.DTYPE_CODE[y_ref.dtype],
                                            L.ptr(acc), L.stream_ptr(y_ref.device)))


def linear_parent(weights: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.linear(x, weights[0])


def mlp_parent(down: torch.Tensor) -> Callable:
    def f(weights, x):
        g = torch.nn.functional.linear(x, weights[0])
        u = torch.nn.functional.linear(x, weights[1])
        return torch.nn.functional.linear(torch.nn.functional.silu(g) * u, down)

    return f


@torch.no_grad()
def gemm_loss_fused(x: torch.Tensor, w_ref: torch.Tensor, w_q: torch.Tensor) -> torch.Tensor:
    """Single-Linear parent: loss[r] = sum_{t,n} (bf16(x w_ref^T) - bf16(x w_q[r]^T))^2 for all stacked variants
    ``w_q [R, N, K]`` in one tcgen05 kernel, outputs never materialised.  bf16 only.  Returns fp32 [R] (sums)."""
    L.require_cuda(x, w_ref, w_q)
    assert x.dtype == torch.bfloat16 and w_ref.dtype == torch.bfloat16 and w_q.dtype == torch.bfloat16
    x, w_ref, w_q = x.contiguous(), w_ref.contiguous(), w_q.contiguous()
    T, K = x.shape
    R, N, _ = w_q.shape
    lib = L.lib()
    ws_bytes = int(lib.b200q_awq_gemm_loss_workspace(T, K, N, R))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    loss = torch.zeros(R, dtype=torch.float32, device=x.device)
    L.check(lib.b200q_awq_gemm_loss(L.ptr(x), T, K, L.ptr(w_ref), L.ptr(w_q), N, R, L.ptr(loss), L.ptr(ws), ws_bytes, L.stream_ptr(x.device)))
    return loss


@torch.no_grad()
def compute_best_scale(x: torch.Tensor, weights: Sequence[torch.Tensor], parent: Callable, args, n_grid: int = 20,
                       duo_scaling: bool = True, process_group=None, fused_linear: bool = False,
                       token_chunk: int = 8192) -> Tuple[torch.Tensor, float, List[float]]:
    """``AWQModifier._compute_best_scale`` for one mapping.

    x        [T_local, K] inputs of the balance layers (this rank's token shard; rows are independent for Linear / MLP
             parents, so samples are concatenated)
    weights  balance-layer weights [N_i, K];  parent(weights, x_chunk) -> parent-module output
    Returns (best_scales fp32 [K] on the CPU like the reference, best_ratio, losses[n_grid]).  Raises if no ratio gives
    a finite loss.  With ``process_group`` the |x| sums / token counts and the loss accumulators are all-reduced (SUM),
    so every rank returns the same argmin."""
    import torch.distributed as dist

    dev = x.device
    K = x.shape[-1]
    x = x.reshape(-1, K)
    dist_on = process_group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    # ---- statistics
    xsum = abs_sum_cols(x)
    cnt = torch.tensor([float(x.shape[0])], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(xsum, op=dist.ReduceOp.SUM, group=process_group)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=process_group)
    x_mean = xsum / cnt.float()
    w_mean = compute_layer_means(weights, args.group_size) if duo_scaling else None
    ratios = [i / n_grid for i in range(n_grid)]
    scales = awq_scales(x_mean, w_mean, ratios, duo_scaling)
    # ---- losses
    acc = torch.zeros(n_grid + 1, dtype=torch.float32, device=dev)  # [n_grid] sums + numel
    if fused_linear and len(weights) == 1:
        wq = torch.empty((n_grid,) + tuple(weights[0].shape), dtype=weights[0].dtype, device=dev)
        for i in range(n_grid):
            scaled_fake_quantize(weights[0], scales[i], args, out=wq[i])
        acc[:n_grid] = gemm_loss_fused(x, weights[0], wq)
        acc[n_grid] = float(x.shape[0] * weights[0].shape[0])
    else:
        wq = [torch.empty_like(w) for w in weights]
        chunks = [x[t0:t0 + token_chunk] for t0 in range(0, x.shape[0], token_chunk)]
        refs = [parent(list(weights), xc) for xc in chunks]
        numel = sum(r.numel() for r in refs)
        for i in range(n_grid):
            for w, o in zip(weights, wq):
                scaled_fake_quantize(w, scales[i], args, out=o)
            for xc, ref in zip(chunks, refs):
                sq_err_accumulate(ref, parent(wq, xc), acc[i:i + 1])
        acc[n_grid] = float(numel)
    if dist_on:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=process_group)
    host = acc.double().cpu()
    losses = (host[:n_grid] / host[n_grid]).tolist()
    best_err, best_i = float("inf"), -1
    for i, v in enumerate(losses):  # first minimum wins (loss < best_error scan)
        if v < best_err:
            best_err, best_i = v, i
    if best_i < 0:
        raise RuntimeError("AWQ: no finite loss for any ratio")
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
