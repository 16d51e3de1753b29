"""MoE calibration: per-expert ``Linear`` modules and the calibrate-all-experts forward.

The reference calls ``oneshot(..., moe_calibrate_all_experts=True)`` (REF:scripts/do_oneshot.py:186; ``calibrate_moe_context`` in
REF:scripts/old_scripts/main_seed-oss-nvfp4.py:77) and explains why in REF:docs/quantization_tips_and_tricks.md:79-98: experts
stored as fused 3-D parameters are invisible to ``targets: ["Linear"]``, and an expert that the router rarely picks sees too few
calibration tokens.  llmcompressor answers both with per-architecture calibration blocks (LLMC modeling/qwen3_moe.py and friends);
their behaviour, restated here structurally instead of per architecture:

  * the fused experts (``gate_up_proj [E, 2I, H]``, ``down_proj [E, H, I]``) become E modules with ``gate_proj`` / ``up_proj`` /
    ``down_proj`` (``w1`` / ``w3`` / ``w2`` for Mixtral / MiniMax) Linears -- the names the recipes' regexes and the compressed
    checkpoint use -- permanent;
  * the BLOCK is kept: its forward, routing (router arguments, correction-bias buffers), shared experts and names are untouched;
    only its ``experts`` child is linearized (fused) or wrapped in place (ModuleList);
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
    """One expert as three Linears: down(act(gate(x)) * up(x)).  ``names`` = the checkpoint's leaf names for (gate, up, down):
    ``gate_proj / up_proj / down_proj`` (Qwen3-MoE, GLM) or ``w1 / w3 / w2`` (Mixtral, MiniMax-M2 -- the reference's MiniMax recipes
    target ``experts.\\d+.(w1|w2|w3)``, REF:configs/recipes/recipe_Minimax-M2.1-Experts-only-AWQ.yaml:20-34)."""

    def __init__(self, gate_up: torch.Tensor, down: torch.Tensor, act_fn, names: Tuple[str, str, str] = ("gate_proj", "up_proj", "down_proj")):
        super().__init__()
        inter2, hidden = gate_up.shape
        inter = inter2 // 2
        self._names = tuple(names)
        setattr(self, names[0], _linear_from(gate_up[:inter]))
        setattr(self, names[1], _linear_from(gate_up[inter:]))
        setattr(self, names[2], _linear_from(down))
        self.act_fn = act_fn

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        g, u, d = (getattr(self, n) for n in self._names)
        return d(self.act_fn(g(x)) * u(x))


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


def expert_leaf_names(block: Optional[torch.nn.Module]) -> Tuple[str, str, str]:
    """(gate, up, down) leaf names of the architecture's per-expert checkpoint tensors."""
    cls = type(block).__name__.lower() if block is not None else ""
    if "mixtral" in cls or "minimax" in cls:
        return ("w1", "w3", "w2")
    return ("gate_proj", "up_proj", "down_proj")


class LinearizedExperts(torch.nn.ModuleList):
    """Drop-in for a transformers-5 fused ``*Experts`` module (3-D ``gate_up_proj [E, 2I, H]`` / ``down_proj [E, H, I]``, called as
    ``experts(hidden_states, top_k_index, top_k_weights)``): the same routed accumulation (``index_add_`` per expert, ascending,
    in the activation dtype) over per-expert ``ExpertMLP`` modules, so ``targets: ["Linear"]`` and the recipes' regexes see the
    experts.  With ``calibrate_all_experts`` every expert runs on ALL tokens (its Linears' hooks see every calibration token)
    but only the routed rows enter the output.  The BLOCK around it -- its forward, router call, buffers such as
    ``e_score_correction_bias``, shared experts -- is untouched."""

    def __init__(self, experts: Iterator[torch.nn.Module], calibrate_all_experts: bool = False):
        super().__init__(list(experts))
        self.calibrate_all_experts = calibrate_all_experts

    def forward(self, hidden_states: torch.Tensor, top_k_index: torch.Tensor, top_k_weights: torch.Tensor) -> torch.Tensor:
        out = torch.zeros_like(hidden_states)
        n_exp = len(self)
        mask = torch.nn.functional.one_hot(top_k_index, num_classes=n_exp).permute(2, 1, 0)  # [E, k, T]
        for e in range(n_exp):
            pos, tok = torch.where(mask[e])
            if self.calibrate_all_experts:
                y = self[e](hidden_states)[tok]
            elif tok.numel():
                y = self[e](hidden_states[tok])
            else:
                continue
            if tok.numel():
                out.index_add_(0, tok, (y * top_k_weights[tok, pos, None]).to(out.dtype))
        return out


def linearize_experts(experts: torch.nn.Module, names: Tuple[str, str, str] = ("gate_proj", "up_proj", "down_proj")) -> torch.nn.ModuleList:
    """Fused 3-D experts -> ``LinearizedExperts`` of ``ExpertMLP`` (a ModuleList of expert modules is returned unchanged)."""
    if _is_expert_list(experts):
        return experts
    if not _is_fused_experts(experts):
        raise ValueError(f"{type(experts).__name__}: neither fused 3-D experts (gate_up_proj / down_proj) nor a ModuleList of experts")
    gu, dn = experts.gate_up_proj.detach(), experts.down_proj.detach()
    if gu.shape[0] != dn.shape[0] or gu.shape[1] != 2 * dn.shape[2] or gu.shape[2] != dn.shape[1]:
        raise ValueError(f"inconsistent fused expert shapes {tuple(gu.shape)} / {tuple(dn.shape)}")
    act = getattr(experts, "act_fn", None) or torch.nn.functional.silu
    return LinearizedExperts(ExpertMLP(gu[e], dn[e], act, names) for e in range(gu.shape[0]))


class _AllTokensExpert:
    """Replacement ``forward`` for ONE expert of a transformers-4 style block (``experts`` is a ModuleList and the BLOCK's own
    forward gathers the routed rows and calls ``experts[e](rows)``).  While ``shared["calibrate"]`` is set, the expert's original
    forward runs ONCE on all tokens of the block input (stashed by a forward pre-hook on the block), so hooks on its Linears see
    every calibration token exactly once, and the rows the block asked for are picked out of that result by matching row contents
    (identical rows give identical outputs, so any match is the right one).  The expert module, its parameters and their names
    stay in place."""

    def __init__(self, orig_forward, shared: dict):
        self.orig = orig_forward
        self._shared = shared

    def __call__(self, x: torch.Tensor, *args, **kwargs):
        sh = self._shared
        if not sh.get("calibrate") or sh.get("x") is None or args or kwargs:
            return self.orig(x, *args, **kwargs)
        x_all = sh["x"]
        cache = sh.setdefault("y", {})
        y_all = cache.get(id(self))
        if y_all is None:                       # once per block forward, however often the block calls this expert
            y_all = cache[id(self)] = self.orig(x_all)
        rows = x.reshape(-1, x.shape[-1])
        if rows.shape[0] == 0:
            return y_all[:0].reshape(*x.shape[:-1], y_all.shape[-1])
        key_all, key = _row_keys(x_all, sh), _row_keys(rows, sh)
        order = torch.argsort(key_all)
        pos = torch.searchsorted(key_all[order], key).clamp(max=order.numel() - 1)
        idx = order[pos]
        if not torch.equal(key_all[idx], key):
            raise RuntimeError("calibrate-all-experts: the block passed rows to an expert that are not rows of the block input")
        return y_all[idx].reshape(*x.shape[:-1], y_all.shape[-1])


def _row_keys(x: torch.Tensor, sh: dict) -> torch.Tensor:
    """64-bit content hash per row (exact on the raw bits; collisions only between rows a 2 x 64-bit weighted sum cannot tell apart)."""
    bits = x.contiguous().view(torch.int16 if x.element_size() == 2 else torch.int32).to(torch.int64)
    w = sh.get("w")
    if w is None or w.numel() != bits.shape[-1] or w.device != bits.device:
        g = torch.Generator(device="cpu").manual_seed(0x5EED)
        w = sh["w"] = torch.randint(-(1 << 40), 1 << 40, (bits.shape[-1],), generator=g, dtype=torch.int64).to(bits.device)
    return (bits * w).sum(-1)


class CalibrationSparseMoeBlock(torch.nn.Module):
    """A stand-alone top-k sparse MoE block (``gate`` router + per-expert modules) with the calibrate-all-experts switch, for
    callers that build their own block (tests, synthetic benchmarks).  ``replace_moe_blocks`` does NOT swap model blocks for this
    class any more: it keeps the model's own block and only linearizes / wraps its ``experts``.

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
    return True


def replace_moe_blocks(model: torch.nn.Module, calibrate_all_experts: bool = True) -> List[str]:
    """Prepare every sparse MoE block of ``model`` (a module with a ``gate`` router and ``experts``) for calibration WITHOUT
    replacing the block: its forward, routing (router arguments, ``e_score_correction_bias`` and other buffers), shared experts
    and parameter names stay exactly as the architecture defines them.  Only ``experts`` changes:

      * fused 3-D experts (transformers 5)      -> ``LinearizedExperts`` with the architecture's per-expert leaf names
                                                   (``gate_proj/up_proj/down_proj``, or ``w1/w3/w2`` for Mixtral / MiniMax)
      * ``ModuleList`` of experts (transformers 4) -> each expert's ``forward`` routed through ``_AllTokensExpert`` in place
                                                   (modules, parameters and names unchanged)

    A block whose ``experts`` is neither raises ``NotImplementedError`` (nothing is skipped silently).  Returns the block names."""
    todo = [(name, m) for name, m in model.named_modules() if _is_sparse_moe_block(m)]
    for name, m in todo:
        ex = m.experts
        if isinstance(ex, LinearizedExperts):
            ex.calibrate_all_experts = calibrate_all_experts
        elif _is_fused_experts(ex):
            new = linearize_experts(ex, expert_leaf_names(m))
            new.calibrate_all_experts = calibrate_all_experts
            m.experts = new
        elif _is_expert_list(ex):
            _wrap_expert_list(m, calibrate_all_experts)
        else:
            raise NotImplementedError(f"{name} ({type(m).__name__}): `experts` is a {type(ex).__name__}, neither fused 3-D gate_up_proj / "
                                      "down_proj parameters nor a ModuleList of expert modules -- cannot linearize it for calibration")
    return [n for n, _ in todo]


def _wrap_expert_list(block: torch.nn.Module, calibrate: bool) -> None:
    """transformers-4 style block: stash the block input for the experts and route every expert call through ``_AllTokensExpert``
    by patching the expert's ``forward`` (the module, its parameters and their names stay in place)."""
    shared = getattr(block, "_b200q_moe_shared", None)
    if shared is None:
        shared = block._b200q_moe_shared = {"calibrate": calibrate, "x": None, "w": None}

        def pre(mod, args, kwargs):
            x = args[0] if args else kwargs.get("hidden_states")
            shared["x"] = x.detach().reshape(-1, x.shape[-1]) if torch.is_tensor(x) else None
            shared["y"] = {}

        def post(mod, args, out):
            shared["x"] = None
            shared["y"] = {}

        block.register_forward_pre_hook(pre, with_kwargs=True)
        block.register_forward_hook(post)
        for e in block.experts:
            e.forward = _AllTokensExpert(e.forward, shared)   # instance attribute: Module.__call__ (and its hooks) go through it
    shared["calibrate"] = calibrate


def _set_calibrate(model: torch.nn.Module, flag: bool) -> None:
    for m in model.modules():
        if isinstance(m, (CalibrationSparseMoeBlock, LinearizedExperts)):
            m.calibrate_all_experts = flag
        sh = getattr(m, "_b200q_moe_shared", None)
        if sh is not None:
            sh["calibrate"] = flag


@contextlib.contextmanager
def moe_calibrate_all_experts(model: torch.nn.Module) -> Iterator[List[str]]:
    """Inside the context every MoE block routes all tokens through all experts (block output unchanged); afterwards the experts
    stay linearized but run sparsely again."""
    names = replace_moe_blocks(model, calibrate_all_experts=True)
    _set_calibrate(model, True)
    try:
        yield names
    finally:
        _set_calibrate(model, False)
