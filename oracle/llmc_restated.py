"""CPU restatement of the llmcompressor half of the hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED for this file: ``llmcompressor >= 0.9`` (/root/reference/pyproject.toml:9, PyPI ``llmcompressor``,
repo vllm-project/llm-compressor) is not installed in this image and has no source on disk, and the reference's
own tests hold no numeric vectors.  The functions below restate the published algorithm as
recalled in SURVEY.md Appendix A, anchored on the reference's call sites
(/root/reference/scripts/do_oneshot.py:179-187, configs/recipes/recipe_awq_w4a16.yaml:13-32) and built on the
PINNED quantization arithmetic (oracle/ct_oracle.c, or live compressed_tensors when available):

  observers/min_max.py     memoryless_minmax / static_minmax / minmax (moving average)       -> MinMaxObserver
  observers/mse.py         shrink-grid MSE observer                                          -> mse_minmax
  modifiers/awq/base.py    _accumulate_mean, _compute_layer_means, _compute_best_scale, _compute_loss, _smooth

Precision conventions where upstream is ambiguous (stated so the CUDA path and this file agree; the per-ratio
loss tolerance of 1e-3 relative covers the alternatives): activation means and the scale vector are fp32; the
smoothing multiply / divide round once to the weight dtype; the parent forward runs in the weight dtype with fp32
accumulation; the loss squares the bf16 difference in fp32.  First minimum wins ties (``loss < best_error``).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from . import oracle as O


# ----------------------------------------------------------------------------- observers (O1-O4)
class MinMaxObserver:
    """memoryless_minmax (no state), static_minmax (running min/max) and minmax (EMA, averaging_constant=0.01)."""

    def __init__(self, kind: str = "memoryless_minmax", averaging_constant: float = 0.01):
        self.kind, self.c = kind, averaging_constant
        self.min = self.max = None

    def update(self, mn: torch.Tensor, mx: torch.Tensor):
        if self.kind == "memoryless_minmax" or self.min is None:
            self.min, self.max = mn.clone(), mx.clone()
        elif self.kind == "static_minmax":
            self.min, self.max = torch.minimum(self.min, mn), torch.maximum(self.max, mx)
        elif self.kind == "minmax":
            if self.c == 1.0:
                self.min, self.max = mn.clone(), mx.clone()
            else:
                self.min = self.min + self.c * (mn - self.min)
                self.max = self.max + self.c * (mx - self.max)
        else:
            raise ValueError(self.kind)
        return self.min, self.max


def activation_global_scale(batches: Sequence[torch.Tensor]) -> torch.Tensor:
    """static_minmax Observer.get_global_scale over calibration batches (NVFP4 input_global_scale)."""
    obs = MinMaxObserver("static_minmax")
    for x in batches:
        if x.numel() == 0:  # calibrate_activations skips empty inputs (unrouted expert)
            continue
        xf = x.reshape(-1)
        obs.update(xf.min().reshape(1), xf.max().reshape(1))
    return O.generate_gparam(float(obs.min), float(obs.max), batches[0].dtype)


def mse_minmax(w: torch.Tensor, geom: O.Geom, qtype: int, num_bits: int, symmetric: bool, maxshrink: float = 0.2,
               patience: int = 5, grid: int = 100, norm: float = 2.4, global_scale: Optional[torch.Tensor] = None):
    """observers/mse.py ``_grid_search_mse``: shrink the (min, max) range on a grid, keep per chunk the range with the smallest
    ``sum |q - x|^norm``; stop after ``patience`` rounds without an improvement in ANY chunk.  GROUP / CHANNEL geometries.

    Dtype chain as upstream's in-place ops on the observed tensor's dtype T: ``p * min`` rounds to T, ``q -= x``, ``abs_``,
    ``pow_(norm)`` each round to T, ``torch.sum`` accumulates in fp32 and rounds to T, ``err < best_error`` compares in T
    (``best_error`` starts at finfo(T).max).  The fake-quantize uses one (scale, zero_point) per chunk -- upstream patches the
    strategy to TOKEN for this; with the chunk's qparams that is the GROUP / CHANNEL fake-quantize of the pinned arithmetic."""
    mn, mx = O.minmax(w, geom)
    best = torch.full(mn.shape, torch.finfo(w.dtype).max, dtype=w.dtype)
    bmn, bmx = mn.clone(), mx.clone()
    R, C = w.shape
    gsz = geom.group if geom.strategy == O.GROUP else C
    no_improve = 0
    for i in range(int(maxshrink * grid)):
        p = 1 - i / grid
        pmn, pmx = p * mn, p * mx
        s, z = O.calculate_qparams(pmn, pmx, qtype, num_bits, symmetric, global_scale)
        q = O.fake_quantize(w, s, z, geom, qtype, num_bits, global_scale)
        q -= w
        q.abs_()
        q.pow_(norm)
        err = torch.sum(q.reshape(R, -1, gsz), dim=-1).reshape(mn.shape)
        better = err < best
        if better.any():
            best[better] = err[better]
            bmn[better], bmx[better] = pmn[better], pmx[better]
            no_improve = 0
        else:
            no_improve += 1
            if no_improve >= patience:
                break
    return bmn, bmx


def accumulate_hessian(x: torch.Tensor, hessian: torch.Tensor, num_samples: int):
    """modifiers/gptq ``accumulate_hessian``: H *= n/(n+t); n += t; inp = sqrt(2/n) * x.float().T; H += inp @ inp.T  (fp32)."""
    import math

    inp = x.reshape(-1, x.shape[-1]).t()
    t = inp.shape[1]
    hessian = hessian * (num_samples / (num_samples + t))
    num_samples += t
    inp = inp.to(torch.float32) * math.sqrt(2 / num_samples)
    return hessian + inp.matmul(inp.t()), num_samples


# ----------------------------------------------------------------------------- AWQ (O5, W1-W5)
def accumulate_abs_mean(batches: Sequence[torch.Tensor]):
    """_accumulate_mean: per-input-channel mean of |x| over all tokens (running sum / count)."""
    total, count = None, 0
    for x in batches:
        x2 = x.reshape(-1, x.shape[-1])
        s = x2.abs().float().sum(0)
        total = s if total is None else total + s
        count += x2.shape[0]
    return total / count, count


def compute_layer_means(weights: Sequence[torch.Tensor], group: int) -> torch.Tensor:
    """_compute_layer_means: mean over all balance-layer rows of |w| / (chunk_absmax + 1e-6), fp64 accumulate."""
    return O.w_mean(list(weights), group).float()


def awq_scales(x_mean: torch.Tensor, w_mean: Optional[torch.Tensor], ratio: float, duo_scaling: bool) -> torch.Tensor:
    if duo_scaling:
        s = (x_mean.pow(ratio) / (w_mean.pow(1 - ratio) + 1e-4)).clamp(min=1e-4)
    else:
        s = x_mean.pow(ratio).clamp(min=1e-4)
    s = s / (s.max() * s.min()).sqrt()
    s[torch.isinf(s)] = 1
    s[torch.isnan(s)] = 1
    return s


def scaled_fake_quantize(w: torch.Tensor, scales: torch.Tensor, geom: O.Geom, qtype: int, num_bits: int, symmetric: bool):
    """W.mul_(s) -> fresh memoryless_minmax observer -> forward_quantize -> / s  (the _compute_best_scale inner step)."""
    ws = (w.float() * scales.view(1, -1)).to(w.dtype)
    mn, mx = O.minmax(ws, geom)
    s, z = O.calculate_qparams(mn, mx, qtype, num_bits, symmetric)
    fq = O.fake_quantize(ws, s, z if qtype == O.INT else torch.zeros(1), geom, qtype, num_bits)
    return (fq.float() / scales.view(1, -1)).to(w.dtype)


def compute_loss(ref: Sequence[torch.Tensor], out: Sequence[torch.Tensor]) -> float:
    """_compute_loss: sum_b sum((fp16_b - int_w_b)^2) / sum_b numel, difference taken in the output dtype."""
    loss, n = 0.0, 0
    for a, b in zip(ref, out):
        loss += float((a - b).float().pow(2).sum())
        n += a.numel()
    return loss / n


def compute_best_scale(x_batches: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                       parent: Callable[[List[torch.Tensor], torch.Tensor], torch.Tensor], geom: O.Geom, qtype: int,
                       num_bits: int, symmetric: bool, n_grid: int = 20, duo_scaling: bool = True):
    """AWQModifier._compute_best_scale for one mapping.

    x_batches: inputs of the balance layers per calibration sample ``[S, K]``; weights: balance-layer weights
    ``[N_i, K]``; parent(weights, x) -> output of the mapping's parent module.  Returns (best_scales, best_ratio,
    losses[n_grid])."""
    x_mean, _ = accumulate_abs_mean(x_batches)
    w_mean = compute_layer_means(weights, geom.group) if duo_scaling else None
    ref = [parent(list(weights), x) for x in x_batches]
    best_err, best_ratio, best_scales, losses = float("inf"), -1, None, []
    for i in range(n_grid):
        ratio = i / n_grid
        s = awq_scales(x_mean, w_mean, ratio, duo_scaling)
        wq = [scaled_fake_quantize(w, s, geom, qtype, num_bits, symmetric) for w in weights]
        out = [parent(wq, x) for x in x_batches]
        loss = compute_loss(ref, out)
        losses.append(loss)
        if loss < best_err:
            best_err, best_ratio, best_scales = loss, ratio, s.clone()
    if best_ratio == -1:
        raise RuntimeError("AWQ: no finite loss for any ratio")
    return best_scales, best_ratio, losses


def smooth(weights: Sequence[torch.Tensor], smooth_weight: torch.Tensor, scales: torch.Tensor):
    """_smooth: balance W *= s; smooth layer (1-D norm weight, or last len(s) rows of a 2-D weight) /= s."""
    new_w = [(w.float() * scales.view(1, -1)).to(w.dtype) for w in weights]
    if smooth_weight.ndim == 1:
        new_s = (smooth_weight.float() / scales).to(smooth_weight.dtype)
    else:
        new_s = smooth_weight.clone()
        k = scales.numel()
        new_s[-k:] = (smooth_weight[-k:].float() / scales.view(-1, 1)).to(smooth_weight.dtype)
    return new_w, new_s


def linear_parent(weights: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.linear(x, weights[0])


def mlp_parent(down: torch.Tensor):
    """parent of gate/up: mlp.forward = down(silu(gate x) * up x) with the un-quantised down_proj in the loop."""

    def f(weights: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
        g = torch.nn.functional.linear(x, weights[0])
        u = torch.nn.functional.linear(x, weights[1])
        return torch.nn.functional.linear(torch.nn.functional.silu(g) * u, down)

    return f


def moe_block_parent(w2: Sequence[torch.Tensor], topk_idx: torch.Tensor, topk_w: torch.Tensor):
    """parent of the layer-wide MoE mapping (post_attention_layernorm -> every expert's w1, w3): the routed sparse-MoE block as
    transformers writes it (Mixtral / Qwen3-MoE style): per expert, gather its routed tokens, ``w2(silu(w1 x) * (w3 x))``, scale
    by the routing weight (cast to the hidden dtype) and ``index_add_`` into the bf16 output, experts in index order.
    ``weights = [w1_0, w3_0, w1_1, w3_1, ...]``; ``x`` holds ALL calibration tokens (routing rows line up with it)."""

    def f(weights: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
        lin = torch.nn.functional.linear
        out = torch.zeros_like(x[:, : w2[0].shape[0]])
        for e in range(len(w2)):
            tok, slot = torch.where(topk_idx == e)
            if tok.numel() == 0:
                continue
            xe = x[tok]
            h = torch.nn.functional.silu(lin(xe, weights[2 * e])) * lin(xe, weights[2 * e + 1])
            y = lin(h, w2[e]) * topk_w[tok, slot, None].to(x.dtype)
            out.index_add_(0, tok, y.to(x.dtype))
        return out

    return f


def attention_parent(o_proj: torch.Tensor, n_heads: int, n_kv: int, head_dim: int, q_norm: torch.Tensor, k_norm: torch.Tensor,
                     rope_theta: float = 1e6, eps: float = 1e-6):
    """parent of q/k/v: transformers ``Qwen3Attention.forward`` on one calibration sample ``x [S, K]`` (batch 1, causal mask,
    positions 0..S-1): q/k per-head RMSNorm, rotary embedding, grouped-query softmax attention, o_proj.  Written with
    elementary ops (explicit softmax) so it does not share code with the product path."""

    def rms(x, w):
        v = x.float()
        v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)
        return w * v.to(x.dtype)

    def f(weights: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
        S = x.shape[0]
        lin = torch.nn.functional.linear
        q = lin(x, weights[0]).view(S, n_heads, head_dim)
        k = lin(x, weights[1]).view(S, n_kv, head_dim)
        v = lin(x, weights[2]).view(S, n_kv, head_dim)
        q, k = rms(q, q_norm).transpose(0, 1), rms(k, k_norm).transpose(0, 1)  # [H, S, d]
        v = v.transpose(0, 1)
        inv = 1.0 / (rope_theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
        fr = torch.outer(torch.arange(S, dtype=torch.float32), inv)
        emb = torch.cat((fr, fr), dim=-1)
        cos, sin = emb.cos().to(x.dtype), emb.sin().to(x.dtype)

        def rot(t):
            t1, t2 = t[..., : head_dim // 2], t[..., head_dim // 2:]
            return t * cos + torch.cat((-t2, t1), dim=-1) * sin

        q, k = rot(q), rot(k)
        rep = n_heads // n_kv
        k, v = k.repeat_interleave(rep, dim=0), v.repeat_interleave(rep, dim=0)
        att = (q.float() @ k.float().transpose(-1, -2)) / (head_dim ** 0.5)
        att = att.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
        p = torch.softmax(att, dim=-1).to(x.dtype)
        o = (p.float() @ v.float()).to(x.dtype).transpose(0, 1).reshape(S, n_heads * head_dim)
        return lin(o, o_proj)

    return f
