"""``oneshot``-shaped driver for the hot path: recipe in, compressed state dict out.

The reference's entry point is ``llmcompressor.oneshot(model, recipe, dataset, ...)`` followed by
``model.save_pretrained(..., save_compressed=True)`` (REF:scripts/do_oneshot.py:179-197).  llmcompressor is not installed in
this image, so -- as SURVEY.md §8b suggests -- this module owns a minimal modifier pair that parses the SAME YAML
blocks (``quantizers_b200.recipe``) and drives the CUDA path directly:

  QuantizationModifier (RTN)  on_start: weight global scales (TENSOR_GROUP) -> fused q/k/v and gate/up minimum
                              (LLMC update_fused_layer_weight_global_scales) -> update_weight_zp_scale for every target;
                              then Compressor.compress per target (CT:compressors/*/base.py)            -> ``quantize_model``
  AWQModifier                 per resolved mapping: _compute_best_scale on the captured balance-layer inputs, _smooth, and the
                              final RTN of its targets (W1-W5)                                            -> ``awq_model``

Both work on any ``torch.nn.Module`` whose target ``nn.Linear`` weights are on a CUDA device; module names, target / ignore
patterns and the emitted state-dict keys (``<module>.weight_packed`` ...) follow compressed-tensors.  Activation capture uses
forward pre-hooks on the balance layers (torch plumbing); the arithmetic is the C ABI's.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import torch

from . import awq as _awq
from . import ops
from .recipe import (ConfigGroup, ModifierSpec, Recipe, default_mappings, group_to_dict, load_recipe, parse_recipe, resolve_mappings,
                     resolve_targets)

FUSED_SIBLINGS = (("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj"))  # LLMC modifiers/utils/helpers.py


def _as_recipe(recipe: Union[str, dict, Recipe]) -> Recipe:
    if isinstance(recipe, Recipe):
        return recipe
    if isinstance(recipe, str) and "\n" not in recipe and recipe.endswith((".yaml", ".yml")):
        return load_recipe(recipe)
    return parse_recipe(recipe)


def _linears(model: torch.nn.Module) -> List[Tuple[str, torch.nn.Module]]:
    return [(n, m) for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)]


def _is_nvfp4(a) -> bool:
    return a.type == "float" and a.num_bits == 4


def _fused_global_scales(targets: Dict[str, ConfigGroup], modules: Dict[str, torch.nn.Module]) -> Dict[str, torch.Tensor]:
    """NVFP4 weight global scales with the fused-sibling minimum: q/k/v of one attention module and gate/up of one MLP share
    min(global_scale) (LLMC update_fused_layer_weight_global_scales; only when ALL siblings of the set are NVFP4 targets)."""
    gs: Dict[str, torch.Tensor] = {}
    for name, g in targets.items():
        if _is_nvfp4(g.weights):
            gs[name] = ops.weight_global_scales(modules[name].weight.detach().unsqueeze(0)).reshape(1)
    parents: Dict[str, List[str]] = {}
    for name in gs:
        parent, _, leaf = name.rpartition(".")
        parents.setdefault(parent, []).append(leaf)
    for parent, leaves in parents.items():
        for sibs in FUSED_SIBLINGS:
            if all(s in leaves for s in sibs):
                names = [f"{parent}.{s}" if parent else s for s in sibs]
                m = torch.stack([gs[n] for n in names]).amin(dim=0)
                for n in names:
                    gs[n] = m
    return gs


@torch.no_grad()
def quantize_model(model: torch.nn.Module, recipe: Union[str, dict, Recipe], modifier: str = "QuantizationModifier",
                   spec: Optional[ModifierSpec] = None) -> Tuple[Dict[str, torch.Tensor], dict]:
    """RTN: observer -> qparams -> quantize -> pack for every Linear the modifier targets; one fused launch per weight.

    Returns (compressed state-dict entries ``{module}.{weight_packed|weight|weight_scale|...}``, the ``quantization_config``
    dict compressed-tensors writes into ``config.json``: format per group, targets, ignore, weights args)."""
    spec = spec or _as_recipe(recipe).modifier(modifier)
    if spec is None:
        raise ValueError(f"the recipe has no {modifier}")
    lin = _linears(model)
    modules = dict(lin)
    targets = resolve_targets(lin, spec)
    for name in targets:
        if not modules[name].weight.is_cuda:
            raise RuntimeError(f"{name}.weight is on {modules[name].weight.device}: the quantization hot path has no CPU fallback")
    gscales = _fused_global_scales(targets, modules)
    sd: Dict[str, torch.Tensor] = {}
    for name, g in targets.items():
        if g.weights is None:
            continue
        w = modules[name].weight.detach()
        out = ops.compress_weight(w, g.weights, global_scale=gscales.get(name))
        for k, v in out.items():
            sd[f"{name}.{k}"] = v
        if modules[name].bias is not None:
            sd[f"{name}.bias"] = modules[name].bias.detach()
    cfg = {"quant_method": "compressed-tensors", "quantization_status": "compressed", "ignore": list(spec.ignore),
           "config_groups": {}}
    formats = set()
    for g in spec.config_groups:
        if g.weights is None:
            continue
        # weights AND input activations (dynamic ones included) + the format compressed-tensors infers from both: a serving
        # engine picks its kernel from this block (FP8_BLOCK = dynamic group-128 fp8 activations, not W8A16)
        cfg["config_groups"][g.name] = group_to_dict(g)
        formats.add(cfg["config_groups"][g.name]["format"])
    cfg["format"] = formats.pop() if len(formats) == 1 else "mixed-precision"
    return sd, cfg


# ----------------------------------------------------------------------------- activation calibration (O3)
@torch.no_grad()
def calibrate_activations(model: torch.nn.Module, spec: ModifierSpec, dataset: Sequence, forward: Optional[Callable] = None) -> Dict[str, torch.Tensor]:
    """LLMC ``calibrate_activations(module, x, "input")`` for every target whose scheme quantizes its inputs statically:
    forward pre-hooks feed each Linear's input to a CUDA observer (state stays on the device, one kernel per call);
      TENSOR_GROUP (NVFP4, ``static_minmax``)  -> ``<module>.input_global_scale`` fp32 [1] = generate_gparam(running min / max)
      TENSOR static (FP8)                      -> ``<module>.input_scale`` = calculate_qparams(min, max) in the activation dtype
    Dynamic schemes (token / group activations) calibrate nothing.  Empty inputs (an expert no token reached) are skipped."""
    from .observers import Observer

    lin = _linears(model)
    targets = resolve_targets(lin, spec)
    observers: Dict[str, Observer] = {}
    handles = []
    for name, mod in lin:
        g = targets.get(name)
        ia = None if g is None else g.input_activations
        if ia is None:
            continue
        strat = ia.strategy
        dyn = getattr(ia, "dynamic", False)
        if not ((strat == "tensor_group") or (strat == "tensor" and dyn in (False, None))):
            continue
        obs = Observer.load_from_registry(getattr(ia, "observer", None) or "static_minmax", base_name="input", args=ia)
        observers[name] = obs

        def hook(m, args, _obs=obs, _tg=(strat == "tensor_group")):
            x = args[0]
            if x.numel() == 0:
                return
            if _tg:
                _obs.get_global_scale(x.detach())
            else:
                _obs(x.detach())

        handles.append(mod.register_forward_pre_hook(hook))
    if not observers:
        return {}
    try:
        for batch in dataset:
            if forward is not None:
                forward(model, batch)
            elif isinstance(batch, dict):
                model(**batch)
            else:
                model(batch)
    finally:
        for h in handles:
            h.remove()
    out: Dict[str, torch.Tensor] = {}
    for name, obs in observers.items():
        if _strategy_of(obs.args) == "tensor_group":
            if obs.global_min is not None:
                out[f"{name}.input_global_scale"] = ops.generate_gparam(obs.global_min, obs.global_max)
        elif obs.min_val is not None:
            scale, _ = ops.calculate_qparams(obs.min_val, obs.max_val, obs.args)
            out[f"{name}.input_scale"] = scale
    return out


def _strategy_of(args) -> str:
    return getattr(args.strategy, "value", args.strategy)


# ----------------------------------------------------------------------------- AWQ over a module tree
class _Capture:
    """Forward pre-hooks that keep the inputs of the balance layers (first positional argument, flattened to [tokens, K]) and
    the positional / keyword arguments of multi-layer parents, per calibration batch -- what LLMC's AWQModifier caches through
    its sequential pipeline (``_setup_activation_cache_hooks``)."""

    def __init__(self, model: torch.nn.Module, layer_names: Iterable[str], parent_names: Iterable[str]):
        self.inputs: Dict[str, List[torch.Tensor]] = {n: [] for n in layer_names}
        self.parent_args: Dict[str, List[Tuple[tuple, dict]]] = {n: [] for n in parent_names}
        self._handles = []
        mods = dict(model.named_modules())
        for n in self.inputs:
            self._handles.append(mods[n].register_forward_pre_hook(self._layer_hook(n)))
        for n in self.parent_args:
            self._handles.append(mods[n].register_forward_pre_hook(self._parent_hook(n), with_kwargs=True))

    def _layer_hook(self, name):
        def hook(mod, args):
            x = args[0].detach()
            self.inputs[name].append(x.reshape(-1, x.shape[-1]))

        return hook

    def _parent_hook(self, name):
        def hook(mod, args, kwargs):
            self.parent_args[name].append((tuple(a.detach() if torch.is_tensor(a) else a for a in args),
                                           {k: (v.detach() if torch.is_tensor(v) else v) for k, v in kwargs.items()}))

        return hook

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def _module_parent(model: torch.nn.Module, parent_name: str, balance_names: Sequence[str], calls: List[Tuple[tuple, dict]]) -> Callable:
    """Generic parent evaluation: run ``parent(*args, **kwargs)`` of every cached call with the balance-layer weights swapped
    for the candidates (torch.func.functional_call); outputs are concatenated over the calls.  ``x`` (the balance-layer input) is
    implied by the cached parent arguments, exactly as upstream's ``_run_samples`` replays ``parent(**kwargs)``."""
    from torch.func import functional_call

    mods = dict(model.named_modules())
    parent = mods[parent_name]
    rel = [b[len(parent_name) + 1:] + ".weight" if parent_name else b + ".weight" for b in balance_names]

    def run(weights: Sequence[torch.Tensor], _x: torch.Tensor) -> torch.Tensor:
        outs = []
        for args, kwargs in calls:
            o = functional_call(parent, {k: w for k, w in zip(rel, weights)}, args, kwargs, strict=False)
            o = o[0] if isinstance(o, (tuple, list)) else o
            outs.append(o.reshape(-1, o.shape[-1]))
        return torch.cat(outs)

    return run


@torch.no_grad()
def awq_model(model: torch.nn.Module, recipe: Union[str, dict, Recipe], calibration: Sequence, forward: Optional[Callable] = None,
              modifier: str = "AWQModifier") -> Tuple[Dict[str, torch.Tensor], dict, Dict[str, Tuple[torch.Tensor, float, List[float]]]]:
    """AWQModifier over a module tree: capture -> per-mapping scale search -> smooth -> RTN of the targets.

    calibration   batches; each is passed to ``forward(model, batch)`` (default ``model(batch)`` / ``model(**batch)``)
    Mappings are resolved from the recipe's ``mappings`` patterns (``resolve_mappings``); a mapping whose balance layers are not
    quantization targets is skipped, as upstream does.  Single-Linear parents run on the fused tensor-core path; multi-layer
    parents replay the parent module with candidate weights (``torch.func.functional_call``) and reduce the squared error with
    ``b200q_sq_err_accumulate``.  Mappings are processed in model order and smoothed before the next capture-dependent mapping of
    the same layer is searched (activations of later mappings are captured once, up front, on the un-smoothed model: smoothing
    is function preserving and quantization is off during calibration).
    Returns (compressed state dict, quantization_config, {smooth_layer -> (best_scales, best_ratio, losses)})."""
    rec = _as_recipe(recipe)
    spec = rec.modifier(modifier)
    if spec is None:
        raise ValueError(f"the recipe has no {modifier}")
    lin = _linears(model)
    targets = resolve_targets(lin, spec)
    names = [n for n, _ in model.named_modules()]
    mods = dict(model.named_modules())
    if not spec.mappings:
        # no `mappings:` in the recipe (REF:configs/recipes/recipe_awq_w4a16.yaml): llmcompressor's model-family defaults
        import dataclasses

        spec = dataclasses.replace(spec, mappings=default_mappings(names))
    resolved = []
    for s_name, balance, parent in resolve_mappings(names, spec):
        balance = [b for b in balance if b in targets]
        sm = mods[s_name]
        if isinstance(sm, torch.nn.Linear) and balance and any(mods[b].in_features != sm.out_features for b in balance):
            # v_proj -> o_proj under grouped-query attention: v has fewer output channels than o has inputs, the scales cannot
            # be folded into v's rows; llmcompressor drops the mapping (modifiers/awq/base.py _set_resolved_mappings)
            continue
        if balance:
            if len(balance) == 1:
                parent = balance[0]
            else:
                # containers have no forward of their own: climb to the first real module (llmcompressor avoids ModuleList
                # ancestors the same way -- for "every expert's w1 / w3" the parent is the sparse-MoE block, not `experts`)
                while parent and isinstance(mods[parent], (torch.nn.ModuleList, torch.nn.ModuleDict)):
                    parent = parent.rpartition(".")[0]
            resolved.append((s_name, balance, parent))
    cap = _Capture(model, {b for _, bl, _ in resolved for b in bl[:1]}, {p for _, bl, p in resolved if len(bl) > 1})
    try:
        for batch in calibration:
            if forward is not None:
                forward(model, batch)
            elif isinstance(batch, dict):
                model(**batch)
            else:
                model(batch)
    finally:
        cap.close()
    results = {}
    for s_name, balance, parent in resolved:
        weights = [mods[b].weight.data for b in balance]
        args = targets[balance[0]].weights
        xs = cap.inputs[balance[0]]
        if not xs:  # e.g. an expert no token was routed to and calibrate-all-experts is off
            continue
        x = torch.cat(xs)
        if len(balance) == 1:
            par = _awq.linear_parent
        else:
            par = _module_parent(model, parent, balance, cap.parent_args[parent])
        # a replayed parent evaluates all cached calls at once: one "chunk" spanning every token
        res = _awq.compute_best_scale(x, weights, par, args, n_grid=spec.n_grid, duo_scaling=bool(spec.duo_scaling),
                                      fused=(len(balance) == 1 and x.dtype == torch.bfloat16), token_chunk=max(int(x.shape[0]), 1))
        smooth_mod = mods[s_name]
        _awq.smooth(weights, smooth_mod.weight.data, res[0])
        if getattr(smooth_mod, "bias", None) is not None:
            smooth_mod.bias.data.copy_((smooth_mod.bias.data.float() / res[0].to(smooth_mod.bias.device)).to(smooth_mod.bias.dtype))
        results[s_name + " -> " + ",".join(balance)] = res
    sd, cfg = quantize_model(model, rec, spec=spec)
    return sd, cfg, results


def oneshot(model: torch.nn.Module, recipe: Union[str, dict, Recipe], dataset: Optional[Sequence] = None,
            forward: Optional[Callable] = None, moe_calibrate_all_experts: bool = False, **_ignored) -> Tuple[Dict[str, torch.Tensor], dict]:
    """``llmcompressor.oneshot``-shaped entry (REF:scripts/do_oneshot.py:179-187): applies every modifier of the recipe in order
    and returns (compressed state dict, quantization_config).  ``dataset`` = calibration batches (needed by AWQModifier).
    ``moe_calibrate_all_experts`` (REF:scripts/do_oneshot.py:186): sparse MoE blocks get per-expert Linears and, while the
    modifiers run, every expert sees every calibration token (``moe_calibration.moe_calibrate_all_experts``)."""
    if moe_calibrate_all_experts:
        from .moe_calibration import moe_calibrate_all_experts as _ctx

        with _ctx(model):
            return oneshot(model, recipe, dataset, forward, moe_calibrate_all_experts=False)
    rec = _as_recipe(recipe)
    sd: Dict[str, torch.Tensor] = {}
    cfgs = []
    for spec in rec.modifiers:
        if spec.kind == "AWQModifier":
            if dataset is None:
                raise ValueError("AWQModifier needs calibration data (dataset=...)")
            part, cfg, _ = awq_model(model, Recipe([spec]), dataset, forward)
        elif spec.kind in ("QuantizationModifier", "GPTQModifier"):
            if spec.kind == "GPTQModifier":
                raise NotImplementedError("GPTQ is outside the hot path this package accelerates (SURVEY.md §8f)")
            part, cfg = quantize_model(model, rec, spec=spec)
            if dataset is not None:
                part.update(calibrate_activations(model, spec, dataset, forward))
        else:
            raise NotImplementedError(f"modifier {spec.kind} is not part of the quantization hot path")
        sd.update(part)
        cfgs.append(cfg)
    cfg = cfgs[0]
    for extra in cfgs[1:]:
        for k, v in extra["config_groups"].items():
            cfg["config_groups"][k if k not in cfg["config_groups"] else f"{k}_{len(cfg['config_groups'])}"] = v
        cfg["ignore"] = sorted(set(cfg["ignore"]) | set(extra["ignore"]))
        if extra["format"] != cfg["format"]:
            cfg["format"] = "mixed-precision"
    return sd, cfg


# ----------------------------------------------------------------------------- save_pretrained(save_compressed=True)
_TORCH_TO_ST = {torch.bfloat16: "BF16", torch.float16: "F16", torch.float32: "F32", torch.float64: "F64", torch.int64: "I64", torch.int32: "I32",
                torch.int16: "I16", torch.int8: "I8", torch.uint8: "U8", torch.bool: "BOOL", torch.float8_e4m3fn: "F8_E4M3",
                torch.float8_e5m2: "F8_E5M2"}


def save_compressed(model: torch.nn.Module, state_dict: Dict[str, torch.Tensor], quantization_config: dict, save_directory: str,
                    config: Optional[dict] = None, max_shard_bytes: int = 5 << 30) -> dict:
    """``model.save_pretrained(save_directory, save_compressed=True)`` (REF:scripts/do_oneshot.py:197) for the result of
    ``oneshot``: the compressed tensors replace the ``weight`` of every quantized module, every other parameter / buffer of the
    model is written as is, shards of at most ``max_shard_bytes`` in safetensors format plus ``model.safetensors.index.json`` and
    ``config.json`` with the ``quantization_config`` block.  Tensors leave the GPU one at a time (D2H into the shard file)."""
    import json
    import os

    from .model_free import build_header

    quantized = {k.rsplit(".", 1)[0] for k in state_dict if k.rsplit(".", 1)[-1] in ("weight_packed", "weight_scale")}
    entries: List[Tuple[str, torch.Tensor]] = []
    for name, t in model.state_dict().items():
        mod, _, leaf = name.rpartition(".")
        if mod in quantized and leaf == "weight":
            continue
        if name in state_dict:
            continue
        entries.append((name, t))
    entries += list(state_dict.items())
    entries.sort(key=lambda kv: kv[0])
    shards: List[List[Tuple[str, torch.Tensor]]] = [[]]
    size = 0
    for name, t in entries:
        nb = t.numel() * t.element_size()
        if shards[-1] and size + nb > max_shard_bytes:
            shards.append([])
            size = 0
        shards[-1].append((name, t))
        size += nb
    os.makedirs(save_directory, exist_ok=True)
    weight_map, total = {}, 0
    for i, shard in enumerate(shards):
        fname = "model.safetensors" if len(shards) == 1 else f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        head, offsets = build_header([(n, _TORCH_TO_ST[t.dtype], tuple(t.shape)) for n, t in shard], {"format": "pt"})
        with open(os.path.join(save_directory, fname), "wb") as f:
            f.write(head)
            for n, t in shard:
                host = t.detach().contiguous().cpu()
                f.write(host.view(torch.uint8).numpy().tobytes() if host.numel() else b"")
                weight_map[n] = fname
                total += host.numel() * host.element_size()
    if len(shards) > 1:
        with open(os.path.join(save_directory, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f, indent=2)
    cfg = dict(config or {})
    cfg["quantization_config"] = quantization_config
    with open(os.path.join(save_directory, "config.json"), "w") as f:
        json.dump(cfg, f, indent=2)
    return {"files": len(shards), "tensors": len(entries), "bytes": total}
