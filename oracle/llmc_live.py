"""The llmcompressor half of the hot path restated on LIVE compressed-tensors calls, device agnostic -- TEST INFRASTRUCTURE ONLY.

``oracle/llmc_restated.py`` runs the same loop on the C oracle (CPU, small shapes).  This module is the same control flow written on
torch ops + the live ``compressed_tensors`` package (``calculate_qparams``, ``fake_quantize``), so that with the tensors on a B200 the
*reference's own arithmetic* evaluates a full BASELINE config-1 decoder layer (T = 64 x 512 tokens, n_grid 20) in seconds: the
full-size AWQ parity test (tests/test_gpu_awq_fullsize.py) and bench.py's CPU baseline / ``--impl reference`` AWQ leg call it.

PARITY UNPINNED for the control flow: llmcompressor (REF:pyproject.toml:9; call site REF:scripts/do_oneshot.py:179-187) is not
installed and has no source on disk; the loop follows SURVEY.md Appendix A lines 488-509.  The arithmetic inside each step IS the
reference's (live CT).  Two places where upstream versions are known to differ are switches (round-1 verdict, item 7):

  x_mean_dtype  "fp32": |x| summed in fp32 over all tokens, divided by the count (this repo's default, order independent);
                "act" : upstream's hook as recalled -- ``x.cpu().abs().flatten(0, -2)`` stays in the activation dtype and the running
                        mean ``(prev_mean * prev_count + inp.sum(0)) / (prev_count + T_b)`` is evaluated in that dtype per batch
  loss_form     "float_pow": ``(a - b).view(-1).float().pow(2).sum()`` per batch (difference in bf16, square + sum in fp32);
                "mse_bf16" : ``F.mse_loss(a, b, reduction="sum")`` per batch on the bf16 tensors (square and result in bf16)
Both accumulate the per-batch value with ``.item()`` into a python float and divide by the total element count.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import torch  # noqa: E402

from . import ct_live as L  # noqa: E402

F = torch.nn.functional


# ----------------------------------------------------------------------------- O5: statistics
def accumulate_abs_mean(batches: Sequence[torch.Tensor], x_mean_dtype: str = "fp32") -> torch.Tensor:
    """LLMC ``_accumulate_mean`` over the calibration batches -> per-input-channel mean of |x| (fp32 [K])."""
    if x_mean_dtype == "fp32":
        total, count = None, 0
        for x in batches:
            x2 = x.reshape(-1, x.shape[-1])
            s = x2.abs().float().sum(0)
            total = s if total is None else total + s
            count += x2.shape[0]
        return total / count
    if x_mean_dtype != "act":
        raise ValueError(x_mean_dtype)
    mean, count = None, 0
    for x in batches:
        inp = x.abs().flatten(0, -2)             # activation dtype
        s = inp.sum(0)                           # torch accumulates in fp32 and rounds the result to the activation dtype
        if mean is None:
            mean = s / inp.shape[0]
        else:
            mean = (mean * count + s) / (count + inp.shape[0])
        count += inp.shape[0]
    return mean.float()


def compute_layer_means(weights: Sequence[torch.Tensor], group_size: Optional[int]) -> torch.Tensor:
    """LLMC ``_compute_layer_means``: mean over all balance-layer rows of |w| / (chunk_absmax + 1e-6), arithmetic in the weight
    dtype, fp64 accumulation (GROUP chunks; ``group_size`` None = one chunk per row)."""
    acc, n = None, 0
    for w in weights:
        org = w.shape
        g = group_size or org[1]
        a = w.abs().reshape(-1, g)
        a = a / (a.amax(dim=1, keepdim=True) + 1e-6)
        a = a.reshape(org)
        s = a.sum(0, dtype=torch.float64)
        acc = s if acc is None else acc + s
        n += org[0]
    return (acc / n).float()


def awq_scales(x_mean: torch.Tensor, w_mean: Optional[torch.Tensor], ratio: float, duo_scaling: bool) -> torch.Tensor:
    if duo_scaling:
        s = (x_mean.pow(ratio) / (w_mean.pow(1 - ratio) + 1e-4)).clamp(min=1e-4)
    else:
        s = x_mean.pow(ratio).clamp(min=1e-4)
    s = s / (s.max() * s.min()).sqrt()
    s[torch.isinf(s)] = 1
    s[torch.isnan(s)] = 1
    return s


# ----------------------------------------------------------------------------- W1 inner step on live CT
def scaled_fake_quantize(w: torch.Tensor, scales: torch.Tensor, args) -> torch.Tensor:
    """``W.mul_(s)`` -> fresh memoryless_minmax observer (live ``calculate_qparams``) -> live ``fake_quantize`` -> ``/ s``."""
    s = scales.to(w.device).view(1, -1)
    ws = (w * s).to(w.dtype)                    # bf16 * fp32 promotes; copied back into the bf16 Parameter
    scale, zp = L.weight_qparams(ws, args)
    fq = L.fake_quantize(ws, scale, zp, args)
    return (fq / s).to(w.dtype)


def batch_loss(a: torch.Tensor, b: torch.Tensor, loss_form: str = "float_pow") -> float:
    if loss_form == "float_pow":
        return (a - b).view(-1).float().pow(2).sum().item()
    if loss_form == "mse_bf16":
        return F.mse_loss(a, b, reduction="sum").item()
    raise ValueError(loss_form)


def compute_best_scale(x_batches: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                       parent: Callable[[List[torch.Tensor], torch.Tensor], torch.Tensor], args, n_grid: int = 20,
                       duo_scaling: bool = True, x_mean_dtype: str = "fp32", loss_form: str = "float_pow", timers: Optional[dict] = None):
    """``AWQModifier._compute_best_scale``.  Returns (best_scales fp32 [K], best_ratio, losses)."""
    import time

    def tick(key, t0):
        if timers is not None:
            if weights[0].is_cuda:
                torch.cuda.synchronize()
            timers[key] = timers.get(key, 0.0) + time.perf_counter() - t0

    t0 = time.perf_counter()
    x_mean = accumulate_abs_mean(x_batches, x_mean_dtype)
    w_mean = compute_layer_means(weights, getattr(args, "group_size", None)) if duo_scaling else None
    tick("stats", t0)
    t0 = time.perf_counter()
    ref = [parent(list(weights), x) for x in x_batches]
    tick("forward", t0)
    numel = sum(r.numel() for r in ref)
    best_err, best_ratio, best_scales, losses = float("inf"), -1, None, []
    for i in range(n_grid):
        ratio = i / n_grid
        t0 = time.perf_counter()
        s = awq_scales(x_mean, w_mean, ratio, duo_scaling)
        wq = [scaled_fake_quantize(w, s, args) for w in weights]
        tick("weights", t0)
        t0 = time.perf_counter()
        loss = 0.0
        for x, r in zip(x_batches, ref):
            loss += batch_loss(r, parent(wq, x), loss_form)
        tick("forward", t0)
        loss /= numel
        losses.append(loss)
        if loss < best_err:
            best_err, best_ratio, best_scales = loss, ratio, s.clone()
    if best_ratio == -1:
        raise RuntimeError("AWQ: no finite loss for any ratio")
    return best_scales, best_ratio, losses


def smooth(weights: Sequence[torch.Tensor], smooth_weight: torch.Tensor, scales: torch.Tensor):
    """``_smooth``: balance W *= s; smooth layer (1-D norm weight, or the last len(s) rows of a 2-D weight) /= s.  In place."""
    s = scales.to(weights[0].device)
    for w in weights:
        w.copy_((w * s.view(1, -1)).to(w.dtype))
    if smooth_weight.ndim == 1:
        smooth_weight.copy_((smooth_weight / s).to(smooth_weight.dtype))
    else:
        k = s.numel()
        smooth_weight[-k:].copy_((smooth_weight[-k:] / s.view(-1, 1)).to(smooth_weight.dtype))


# ----------------------------------------------------------------------------- parents (W2), elementary torch ops
def linear_parent(weights: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, weights[0])


def mlp_parent(down: torch.Tensor):
    def f(weights: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
        return F.linear(F.silu(F.linear(x, weights[0])) * F.linear(x, weights[1]), down)

    return f


def attention_parent(o_proj: torch.Tensor, n_heads: int, n_kv: int, head_dim: int, q_norm: torch.Tensor, k_norm: torch.Tensor,
                     rope_theta: float = 1e6, eps: float = 1e-6, attn_implementation: str = "sdpa"):
    """transformers ``Qwen3Attention.forward`` on one calibration sample ``x [S, K]`` (batch 1, causal, positions 0..S-1).
    ``attn_implementation``: "sdpa" = transformers' default (``sdpa_attention_forward`` -> ``F.scaled_dot_product_attention`` with
    ``is_causal``), what ``AutoModelForCausalLM.from_pretrained`` in REF:scripts/do_oneshot.py:82-96 gives; "eager" =
    ``eager_attention_forward`` (explicit fp32 softmax, probabilities cast to the activation dtype before P @ V)."""
    dev = o_proj.device

    def rms(x, w):
        v = x.float()
        v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)
        return w * v.to(x.dtype)

    def f(weights: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
        S = x.shape[0]
        q = F.linear(x, weights[0]).view(S, n_heads, head_dim)
        k = F.linear(x, weights[1]).view(S, n_kv, head_dim)
        v = F.linear(x, weights[2]).view(S, n_kv, head_dim)
        q, k = rms(q, q_norm).transpose(0, 1), rms(k, k_norm).transpose(0, 1)  # [H, S, d]
        v = v.transpose(0, 1)
        inv = 1.0 / (rope_theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=dev) / head_dim))
        fr = torch.outer(torch.arange(S, dtype=torch.float32, device=dev), inv)
        emb = torch.cat((fr, fr), dim=-1)
        cos, sin = emb.cos().to(x.dtype), emb.sin().to(x.dtype)

        def rot(t):
            t1, t2 = t[..., : head_dim // 2], t[..., head_dim // 2:]
            return t * cos + torch.cat((-t2, t1), dim=-1) * sin

        q, k = rot(q), rot(k)
        rep = n_heads // n_kv
        k, v = k.repeat_interleave(rep, dim=0), v.repeat_interleave(rep, dim=0)          # repeat_kv
        if attn_implementation == "sdpa":
            o = F.scaled_dot_product_attention(q[None], k[None], v[None], is_causal=True)[0]
            return F.linear(o.transpose(0, 1).reshape(S, n_heads * head_dim), o_proj)
        att = torch.matmul(q, k.transpose(-1, -2)) * (head_dim ** -0.5)
        att = att + torch.full((S, S), float("-inf"), device=dev, dtype=att.dtype).triu(1)
        p = torch.softmax(att, dim=-1, dtype=torch.float32).to(x.dtype)
        o = torch.matmul(p, v).transpose(0, 1).reshape(S, n_heads * head_dim)
        return F.linear(o, o_proj)

    return f


def search_decoder_layer(weights: dict, acts: dict, args, n_heads: int, n_kv: int, head_dim: int, seq_len: int, n_grid: int = 20,
                         duo_scaling: bool = True, apply: bool = True, x_mean_dtype: str = "fp32", loss_form: str = "float_pow",
                         timers: Optional[dict] = None) -> dict:
    """The default Llama/Qwen3 mappings of one dense decoder layer (LLMC modifiers/awq/mappings.py),
    sample by sample (``seq_len`` tokens per calibration batch), smoothing applied between mappings like the sequential pipeline."""
    w = weights
    kw = dict(n_grid=n_grid, duo_scaling=duo_scaling, x_mean_dtype=x_mean_dtype, loss_form=loss_form, timers=timers)

    def batches(x):
        return [x[t0:t0 + seq_len] for t0 in range(0, x.shape[0], seq_len)]

    out = {}
    attn = attention_parent(w["o"], n_heads, n_kv, head_dim, w["q_norm"], w["k_norm"])
    out["qkv"] = compute_best_scale(batches(acts["attn_in"]), [w["q"], w["k"], w["v"]], attn, args, **kw)
    if apply:
        smooth([w["q"], w["k"], w["v"]], w["input_layernorm"], out["qkv"][0])
    if w["v"].shape[0] == w["o"].shape[1] and "o_in" in acts:
        out["v_o"] = compute_best_scale(batches(acts["o_in"]), [w["o"]], linear_parent, args, **kw)
        if apply:
            smooth([w["o"]], w["v"], out["v_o"][0])
    out["gate_up"] = compute_best_scale(batches(acts["mlp_in"]), [w["gate"], w["up"]], mlp_parent(w["down"]), args, **kw)
    if apply:
        smooth([w["gate"], w["up"]], w["post_attention_layernorm"], out["gate_up"][0])
    out["down"] = compute_best_scale(batches(acts["down_in"]), [w["down"]], linear_parent, args, **kw)
    if apply:
        smooth([w["down"]], w["up"], out["down"][0])
    return out
