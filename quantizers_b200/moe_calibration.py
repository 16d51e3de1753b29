"""MoE calibration: per-expert ``Linear`` modules and the calibrate-all-experts forward.

The reference calls ``oneshot(..., moe_calibrate_all_experts=True)`` (REF:scripts/do_oneshot.py:186; ``calibrate_moe_context`` in
REF:scripts/old_scripts/main_seed-oss-nvfp4.py:77) and explains why in REF:docs/quantization_tips_and_tricks.md:79-98: experts
stored as fused 3-D parameters are invisible to ``targets: ["Linear"]``, and an expert that the router rarely picks sees too few
calibration tokens.  llmcompressor answers both with per-architecture calibration blocks (LLMC modeling/qwen3_moe.py and friends);
their behaviour, restated here structurally instead of per architecture:

  * the fused experts (``gate_up_proj [E, 2I, H]``, ``down_proj [E, H, I]``) become E modules with ``gate_proj`` / ``up_proj`` /
    ``down_proj`` Linears (names the recipes' regexes and the compressed checkpoint use) -- permanent;
  * routing is untouched (the block's own router module is called as is);
  * with ``calibrate_all_experts`` every expert runs on ALL tokens -- so the hooks on its Linears (AWQ capture, activation
    observers) see every calibration token -- but only the rows the router selected enter the block output, weighted and
    accumulated expert by expert exactly like the sparse forward, so the hidden states that flow on are unchanged.

This is host-side module plumbing (torch ops; the GEMMs are plain library calls on calibration data, not a hot path).  The AWQ
search itself runs on ``awq.search_expert_mappings`` / ``search_moe_block_mapping`` with what these hooks captured.
"""
from __future__ import annotations

import contextlib
from typing import Iterator, List, Optional, Tuple

import torch


class ExpertMLP(torch.nn.Module):
    """One expert as three Linears: down_proj(act(gate_proj(x)) * up_proj(x))."""

    def __init__(self, gate_up: torch.Tensor, down: torch.Tensor, act_fn):
        super().__init__()
        inter2, hidden = gate_up.shape
        inter = inter2 // 2
        mk = lambda w: _linear_from(w)
        self.gate_proj = mk(gate_up[:inter])
        self.up_proj = mk(gate_up[inter:])
        self.down_proj = mk(down)
        self.act_fn = act_fn

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.down_proj(self.act_fn(self.gate_proj(x)) * self.up_proj(x))


def _linear_from(w: torch.Tensor) -> torch.nn.Linear:
    out_f, in_f = w.shape
    lin = torch.nn.Linear(in_f, out_f, bias=False, device="meta", dtype=w.dtype)
    lin.weight = torch.nn.Parameter(w.detach().clone().contiguous(), requires_grad=False)
    return lin


def _is_fused_experts(m: torch.nn.Module) -> bool:
    gu, dn = getattr(m, "gate_up_proj", None), getattr(m, "down_proj", None)
    return isinstance(gu, torch.Tensor) and isinstance(dn, torch.Tensor) and gu.ndim == 3 and dn.ndim == 3


def _is_expert_list(m: torch.nn.Module) -> bool:
    return isinstance(m, torch.nn.ModuleList) and len(m) > 0 and all(callable(getattr(e, "forward", None)) for e in m)


def linearize_experts(experts: torch.nn.Module) -> torch.nn.ModuleList:
    """Fused 3-D experts -> ModuleList of ``ExpertMLP`` (a ModuleList of expert modules is returned unchanged)."""
    if _is_expert_list(experts):
        return experts
    if not _is_fused_experts(experts):
        raise ValueError(f"{type(experts).__name__}: neither fused 3-D experts (gate_up_proj / down_proj) nor a ModuleList of experts")
    gu, dn = experts.gate_up_proj.detach(), experts.down_proj.detach()
    if gu.shape[0] != dn.shape[0] or gu.shape[1] != 2 * dn.shape[2] or gu.shape[2] != dn.shape[1]:
        raise ValueError(f"inconsistent fused expert shapes {tuple(gu.shape)} / {tuple(dn.shape)}")
    act = getattr(experts, "act_fn", None) or torch.nn.functional.silu
    return torch.nn.ModuleList([ExpertMLP(gu[e], dn[e], act) for e in range(gu.shape[0])])


class CalibrationSparseMoeBlock(torch.nn.Module):
    """Drop-in for a top-k sparse MoE block (``gate`` router + ``experts``) with the calibrate-all-experts switch.

    Router conventions handled: a module returning ``(logits, scores [T, k], indices [T, k])`` (transformers 5 ``*TopKRouter``),
    or a plain ``Linear`` producing logits (transformers 4: softmax in fp32 -> top-k -> optional renormalisation, the block's
    ``top_k`` / ``norm_topk_prob``)."""

    def __init__(self, gate: torch.nn.Module, experts: torch.nn.ModuleList, top_k: Optional[int] = None, norm_topk_prob: bool = True,
                 calibrate_all_experts: bool = True, returns_router_logits: bool = False):
        super().__init__()
        self.gate = gate
        self.experts = experts
        self.top_k = top_k
        self.norm_topk_prob = norm_topk_prob
        self.calibrate_all_experts = calibrate_all_experts
        self.returns_router_logits = returns_router_logits

    def route(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        out = self.gate(x)
        if isinstance(out, tuple):
            logits, scores, idx = out
            return logits, scores, idx
        if self.top_k is None:
            raise ValueError("a Linear router needs top_k")
        probs = torch.nn.functional.softmax(out, dim=-1, dtype=torch.float)
        scores, idx = torch.topk(probs, self.top_k, dim=-1)
        if self.norm_topk_prob:
            scores = scores / scores.sum(dim=-1, keepdim=True)
        return out, scores.to(x.dtype), idx

    def forward(self, hidden_states: torch.Tensor):
        shape = hidden_states.shape
        x = hidden_states.reshape(-1, shape[-1])
        logits, scores, idx = self.route(x)
        out = torch.zeros_like(x)
        n_exp = len(self.experts)
        mask = torch.nn.functional.one_hot(idx, num_classes=n_exp).permute(2, 1, 0)  # [E, k, T]
        for e in range(n_exp):
            pos, tok = torch.where(mask[e])
            if self.calibrate_all_experts:
                y = self.experts[e](x)[tok]
            elif tok.numel():
                y = self.experts[e](x[tok])
            else:
                continue
            if tok.numel():
                out.index_add_(0, tok, (y * scores[tok, pos, None]).to(out.dtype))
        out = out.reshape(shape)
        return (out, logits) if self.returns_router_logits else out


def _is_sparse_moe_block(m: torch.nn.Module) -> bool:
    if isinstance(m, CalibrationSparseMoeBlock):
        return False
    gate, experts = getattr(m, "gate", None), getattr(m, "experts", None)
    if not isinstance(gate, torch.nn.Module) or not isinstance(experts, torch.nn.Module):
        return False
    return _is_fused_experts(experts) or _is_expert_list(experts)


def replace_moe_blocks(model: torch.nn.Module, calibrate_all_experts: bool = True) -> List[str]:
    """Replace every sparse MoE block of ``model`` (a module with a ``gate`` router and ``experts``) by a
    ``CalibrationSparseMoeBlock`` over per-expert Linears.  Blocks with extra trainable parts (shared experts ...) are left alone
    unless they only have ``gate`` and ``experts`` children.  Returns the replaced module names."""
    todo = []
    for name, m in model.named_modules():
        if _is_sparse_moe_block(m):
            extra = [n for n, _ in m.named_children() if n not in ("gate", "experts")]
            if extra:
                continue
            todo.append((name, m))
    for name, m in todo:
        gate = m.gate
        top_k = getattr(m, "top_k", None) or getattr(gate, "top_k", None)
        norm = getattr(m, "norm_topk_prob", getattr(gate, "norm_topk_prob", True))
        new = CalibrationSparseMoeBlock(gate, linearize_experts(m.experts), top_k=top_k, norm_topk_prob=bool(norm),
                                        calibrate_all_experts=calibrate_all_experts)
        parent_name, _, leaf = name.rpartition(".")
        setattr(model.get_submodule(parent_name) if parent_name else model, leaf, new)
    return [n for n, _ in todo]


@contextlib.contextmanager
def moe_calibrate_all_experts(model: torch.nn.Module) -> Iterator[List[str]]:
    """Inside the context every MoE block routes all tokens through all experts (block output unchanged); afterwards the blocks
    stay linearized but run sparsely again."""
    names = replace_moe_blocks(model, calibrate_all_experts=True)
    blocks = [m for m in model.modules() if isinstance(m, CalibrationSparseMoeBlock)]
    for b in blocks:
        b.calibrate_all_experts = True
    try:
        yield names
    finally:
        for b in blocks:
            b.calibrate_all_experts = False
