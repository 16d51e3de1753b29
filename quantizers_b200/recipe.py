"""Recipe (YAML) -> quantization plan, for the schema the reference feeds to ``llmcompressor.oneshot``.

The reference keeps its quantization settings in YAML recipes (REF:configs/recipes/*.yaml, selected by
REF:scripts/do_oneshot.py:150-187) written in llmcompressor's two layouts:

    modifiers:                       |   <stage>:                          (e.g. default_stage / quant_stage)
      - name: AWQModifier            |     <group>_modifiers:              (e.g. default_modifiers / quant_modifiers)
        config_groups: {...}         |       AWQModifier: {...}
        ignore: [...]                |       QuantizationModifier: {...}

A modifier block carries ``targets`` / ``scheme`` (a compressed-tensors preset name, CT:quantization/quant_scheme.py:143-428)
or ``config_groups: {name: {targets, weights: QuantizationArgs, input_activations}}``, ``ignore``, and for AWQ ``mappings``
(``smooth_layer`` / ``balance_layers`` patterns), ``duo_scaling``, ``n_grid``.  This module parses that schema into plain
objects and resolves which module gets which ``QuantizationArgs``, with compressed-tensors' matching rules
(CT:utils/match.py ``is_match``: a target is a module class name, an exact module name, or ``re:<regex>`` matched with
``re.match`` against the module name; ``ignore`` wins).  It is host logic only -- no arithmetic happens here.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

from .scheduler import SchemeArgs

# weight halves of the compressed-tensors preset schemes (CT:quantization/quant_scheme.py:143-428), verified against the
# installed package in tests/test_recipe.py
_PRESET_WEIGHTS: Dict[str, Optional[dict]] = {
    "UNQUANTIZED": None,
    "W8A16": dict(num_bits=8, type="int", symmetric=True, strategy="channel"),
    "W4A16": dict(num_bits=4, type="int", symmetric=True, strategy="group", group_size=128),
    "W4A16_ASYM": dict(num_bits=4, type="int", symmetric=False, strategy="group", group_size=128),
    "W8A8": dict(num_bits=8, type="int", symmetric=True, strategy="channel"),
    "INT8": dict(num_bits=8, type="int", symmetric=True, strategy="channel"),
    "W4A8": dict(num_bits=4, type="int", symmetric=True, strategy="group", group_size=128),
    "W4AFP8": dict(num_bits=4, type="int", symmetric=True, strategy="group", group_size=128),
    "FP8": dict(num_bits=8, type="float", symmetric=True, strategy="tensor"),
    "FP8_DYNAMIC": dict(num_bits=8, type="float", symmetric=True, strategy="channel"),
    "FP8_BLOCK": dict(num_bits=8, type="float", symmetric=True, strategy="block", block_structure=[128, 128]),
    "NVFP4A16": dict(num_bits=4, type="float", symmetric=True, strategy="tensor_group", group_size=16),
    "NVFP4": dict(num_bits=4, type="float", symmetric=True, strategy="tensor_group", group_size=16),
}
# input-activation halves of the presets (CT:quantization/quant_scheme.py:159-428).  Static ones (FP8, NVFP4's global scale) are
# calibrated; dynamic ones carry no observer and calibrate nothing, but they ARE part of the scheme a serving engine reads from
# ``quantization_config`` (FP8_BLOCK = W8A8 with dynamic per-token-group-128 fp8 activations, not W8A16) and they decide the
# compression format (CT:compressors/format.py: INT weights + activations -> int-quantized, not pack-quantized).
_PRESET_INPUTS: Dict[str, dict] = {
    "W8A8": dict(num_bits=8, type="int", symmetric=True, strategy="token", dynamic=True, observer=None),
    "INT8": dict(num_bits=8, type="int", symmetric=True, strategy="token", dynamic=True, observer=None),
    "W4A8": dict(num_bits=8, type="int", symmetric=True, strategy="token", dynamic=True, observer=None),
    "W4AFP8": dict(num_bits=8, type="float", symmetric=True, strategy="token", dynamic=True, observer=None),
    "FP8": dict(num_bits=8, type="float", symmetric=True, strategy="tensor", dynamic=False, observer="memoryless_minmax"),
    "FP8_DYNAMIC": dict(num_bits=8, type="float", symmetric=True, strategy="token", dynamic=True, observer=None),
    "FP8_BLOCK": dict(num_bits=8, type="float", symmetric=True, strategy="group", group_size=128, dynamic=True, observer=None),
    "NVFP4": dict(num_bits=4, type="float", symmetric=True, strategy="tensor_group", group_size=16, dynamic="local", observer="static_minmax"),
}


class RecipeError(ValueError):
    pass


def preset_input_args(name: str) -> Optional[SchemeArgs]:
    """``QuantizationArgs`` of a preset's input activations (None for weight-only presets)."""
    kw = _PRESET_INPUTS.get(str(name).upper())
    return None if kw is None else _args_from_dict(kw)


def infer_format(weights: Optional[SchemeArgs], input_activations: Optional[SchemeArgs] = None) -> str:
    """The compression format compressed-tensors infers for a Linear with this scheme: the first compressor of
    COMPRESSION_FORMAT_PRIORITY whose ``can_compress`` accepts it (CT:compressors/format.py:18-27,75-96; nvfp4/base.py:112-120,
    pack_quantized/base.py:122-130, naive_quantized/base.py:114-149)."""
    if weights is None:
        return "dense"
    if weights.type == "float" and weights.num_bits == 4 and weights.group_size == 16:
        return "nvfp4-pack-quantized"
    if weights.type == "int" and weights.num_bits in (4, 8) and input_activations is None:
        return "pack-quantized"
    if weights.type == "int" and input_activations is not None:
        return "int-quantized"
    if weights.type == "float" and input_activations is not None:
        return "float-quantized"
    return "naive-quantized"


def args_to_dict(a: Optional[SchemeArgs]) -> Optional[dict]:
    """A ``QuantizationArgs`` block as compressed-tensors serialises it into ``quantization_config`` (the fields of
    ``QuantizationArgs.model_dump()``, CT:quantization/quant_args.py:157-262)."""
    if a is None:
        return None
    dyn = getattr(a, "dynamic", False)
    return {"num_bits": a.num_bits, "type": a.type, "symmetric": a.symmetric, "group_size": a.group_size, "strategy": a.strategy,
            "block_structure": a.block_structure, "dynamic": dyn, "actorder": None, "scale_dtype": None, "zp_dtype": None,
            "observer": getattr(a, "observer", "memoryless_minmax"), "observer_kwargs": dict(getattr(a, "observer_kwargs", {}) or {})}


def group_to_dict(g: "ConfigGroup") -> dict:
    """One ``config_groups`` entry of ``quantization_config`` (CT QuantizationScheme.model_dump + the inferred format)."""
    return {"targets": list(g.targets), "weights": args_to_dict(g.weights), "input_activations": args_to_dict(g.input_activations),
            "output_activations": None, "format": infer_format(g.weights, g.input_activations)}


def preset_args(name: str) -> Optional[SchemeArgs]:
    """``QuantizationArgs`` of a preset's weights (None for UNQUANTIZED)."""
    key = str(name).upper()
    if key not in _PRESET_WEIGHTS:
        raise RecipeError(f"unknown preset scheme {name!r}; known: {sorted(_PRESET_WEIGHTS)}")
    kw = _PRESET_WEIGHTS[key]
    return None if kw is None else _args_from_dict(kw)


def _args_from_dict(d: dict) -> SchemeArgs:
    """QuantizationArgs block of a recipe -> SchemeArgs, with the validation CT's pydantic model applies
    (CT:quantization/quant_args.py:263-408): strategy inferred from group_size when absent, group strategies need a group_size."""
    d = dict(d)
    num_bits = int(d.get("num_bits", 8))
    qtype = str(d.get("type", "int")).lower()
    symmetric = bool(d.get("symmetric", True))
    group_size = d.get("group_size")
    strategy = d.get("strategy")
    block = d.get("block_structure")
    if isinstance(block, str):  # "128x128"
        block = [int(v) for v in block.lower().split("x")]
    if strategy is None:
        if group_size is None or int(group_size) == -1:
            strategy = "channel" if group_size is not None else "tensor"
        else:
            strategy = "group"
    strategy = str(strategy).lower()
    if strategy in ("group", "tensor_group"):
        if group_size is None or int(group_size) <= 0:
            raise RecipeError(f"strategy {strategy} requires group_size to be set to a positive value")
        group_size = int(group_size)
    else:
        group_size = None
    if strategy == "block" and block is None:
        raise RecipeError("strategy block requires block_structure")
    if qtype not in ("int", "float"):
        raise RecipeError(f"unknown quantization type {qtype!r}")
    if d.get("actorder") not in (None, False, "none"):
        raise RecipeError("actorder (g_idx) is not supported by the fused kernels (no reference recipe enables it)")
    a = SchemeArgs(num_bits, qtype, symmetric, strategy, group_size, list(block) if block else None)
    a.dynamic = d.get("dynamic", False)
    # CT drops the observer of dynamic (non-local) schemes (quant_args.py validate_model_after)
    a.observer = d["observer"] if "observer" in d else (None if a.dynamic is True else "memoryless_minmax")
    a.observer_kwargs = dict(d.get("observer_kwargs") or {})
    return a


def _clean_pattern(t) -> str:
    """Recipes in the wild carry python-literal debris (``r"re:..."`` inside YAML, REF:configs/recipes/recipe_mixed_fp8_int4.yaml:9)."""
    t = str(t).strip()
    m = re.fullmatch(r"r?([\"'])(.*)\1", t)
    return m.group(2) if m else t


def _as_list(v) -> List[str]:
    if v is None:
        return []
    if isinstance(v, (str, bytes)):
        return [_clean_pattern(v)]
    return [_clean_pattern(x) for x in v]


@dataclass
class ConfigGroup:
    name: str
    targets: List[str]
    weights: Optional[SchemeArgs]
    input_activations: Optional[SchemeArgs] = None
    preset: Optional[str] = None


@dataclass
class Mapping:
    smooth_layer: str
    balance_layers: List[str]


@dataclass
class ModifierSpec:
    kind: str                               # "QuantizationModifier" | "AWQModifier" | ...
    config_groups: List[ConfigGroup] = field(default_factory=list)
    ignore: List[str] = field(default_factory=list)
    mappings: List[Mapping] = field(default_factory=list)
    duo_scaling: Union[bool, str] = True
    n_grid: int = 20
    extra: dict = field(default_factory=dict)


@dataclass
class Recipe:
    modifiers: List[ModifierSpec]

    def modifier(self, kind: str) -> Optional[ModifierSpec]:
        for m in self.modifiers:
            if m.kind == kind:
                return m
        return None


_KNOWN_KEYS = {"targets", "scheme", "config_groups", "ignore", "mappings", "duo_scaling", "n_grid", "name", "group", "offload_device",
               "sequential_targets", "kv_cache_scheme"}


def _parse_modifier(kind: str, body: dict) -> ModifierSpec:
    body = dict(body or {})
    spec = ModifierSpec(kind=kind, ignore=_as_list(body.get("ignore")), duo_scaling=body.get("duo_scaling", True),
                        n_grid=int(body.get("n_grid", 20)))
    groups = body.get("config_groups")
    if groups is None:
        # the reference's mixed recipe names the config-group dict by its content (``mlp_experts_projections:`` directly under the
        # modifier, REF:configs/recipes/recipe_mixed_fp8_int4.yaml:12-13): accept any unknown dict-of-groups key
        for k, v in body.items():
            if k not in _KNOWN_KEYS and isinstance(v, dict) and v and all(isinstance(g, dict) and ("weights" in g or "targets" in g) for g in v.values()):
                groups = v
                break
    if groups:
        for gname, g in groups.items():
            w = g.get("weights")
            ia = g.get("input_activations")
            spec.config_groups.append(ConfigGroup(gname, _as_list(g.get("targets", ["Linear"])), _args_from_dict(w) if w else None,
                                                  _args_from_dict(ia) if isinstance(ia, dict) else None))
    if body.get("scheme") is not None:
        scheme = body["scheme"]
        targets = _as_list(body.get("targets", ["Linear"]))
        if isinstance(scheme, dict):  # {preset: [targets]} form
            for pname, tg in scheme.items():
                spec.config_groups.append(ConfigGroup(f"group_{len(spec.config_groups)}", _as_list(tg), preset_args(pname),
                                                      preset_input_args(pname), preset=str(pname).upper()))
        else:
            key = str(scheme).upper()
            spec.config_groups.append(ConfigGroup(f"group_{len(spec.config_groups)}", targets, preset_args(key),
                                                  preset_input_args(key), preset=key))
    for m in body.get("mappings") or []:
        spec.mappings.append(Mapping(_clean_pattern(m["smooth_layer"]), _as_list(m["balance_layers"])))
    spec.extra = {k: v for k, v in body.items() if k in ("offload_device", "sequential_targets", "kv_cache_scheme")}
    return spec


def parse_recipe(doc: Union[str, dict]) -> Recipe:
    """Parse a recipe given as YAML text or as the already loaded mapping."""
    if isinstance(doc, str):
        import yaml

        doc = yaml.safe_load(doc)
    if not isinstance(doc, dict):
        raise RecipeError("a recipe is a YAML mapping")
    mods: List[ModifierSpec] = []
    if isinstance(doc.get("modifiers"), list):
        for m in doc["modifiers"]:
            kind = m.get("name") or m.get("type")
            if not kind:
                raise RecipeError("entries of `modifiers:` need a `name:`")
            mods.append(_parse_modifier(kind, m))
    for stage, body in doc.items():
        if not isinstance(body, dict):
            continue
        for gname, group in body.items():
            if str(gname).endswith("_modifiers") and isinstance(group, dict):
                for kind, mbody in group.items():
                    mods.append(_parse_modifier(kind, mbody))
    if not mods:
        raise RecipeError("no modifiers found in the recipe")
    return Recipe(mods)


def load_recipe(path: str) -> Recipe:
    with open(path, "r", encoding="utf-8") as f:
        return parse_recipe(f.read())


# ----------------------------------------------------------------------------- target resolution (CT:utils/match.py)
def match_name(name: str, target: str) -> bool:
    if target.startswith("re:"):
        return re.match(target[3:], name) is not None
    return target == name


def _match_class(module, target: str) -> bool:
    return any(c.__name__ == target or (c.__name__ == "LinearBase" and target == "Linear") for c in type(module).__mro__)


def is_match(name: str, module, targets: Iterable[str], ignore: Iterable[str] = ()) -> bool:
    return any(match_name(name, t) or _match_class(module, t) for t in targets) and not any(
        match_name(name, i) or _match_class(module, i) for i in ignore)


def resolve_targets(named_modules: Iterable[Tuple[str, object]], spec: ModifierSpec) -> Dict[str, ConfigGroup]:
    """module name -> config group of this modifier, with compressed-tensors' priority (``match_targets``, CT:utils/match.py:
    targets sorted so that exact names come before ``re:`` patterns, NAME matches tried before CLASS matches, first hit wins) --
    independent of the order the groups are listed in: a ``re:.*mlp.*`` group takes a module from a ``Linear`` group listed first."""
    flat = sorted(((t, g) for g in spec.config_groups for t in g.targets), key=lambda tg: ("re:" in tg[0], tg[0]))
    out: Dict[str, ConfigGroup] = {}
    for name, mod in named_modules:
        if any(match_name(name, i) or _match_class(mod, i) for i in spec.ignore):
            continue
        hit = next((g for t, g in flat if match_name(name, t)), None)
        if hit is None:
            hit = next((g for t, g in flat if _match_class(mod, t)), None)
        if hit is not None:
            out[name] = hit
    return out


# llmcompressor's model-family defaults (modifiers/awq/mappings.py, as recalled in SURVEY.md Appendix A line 508): used when the
# AWQModifier block has no ``mappings:`` -- REF:configs/recipes/recipe_awq_w4a16.yaml (the recipe behind
# REF:configs/test-quantize_qwen3-4b-awq.yaml) relies on them.
DEFAULT_MAPPINGS = [
    ("re:.*input_layernorm$", ["re:.*q_proj$", "re:.*k_proj$", "re:.*v_proj$"]),
    ("re:.*v_proj$", ["re:.*o_proj$"]),
    ("re:.*post_attention_layernorm$", ["re:.*gate_proj$", "re:.*up_proj$"]),
    ("re:.*up_proj$", ["re:.*down_proj$"]),
]
MOE_DEFAULT_MAPPINGS = [
    ("re:.*input_layernorm$", ["re:.*q_proj$", "re:.*k_proj$", "re:.*v_proj$"]),
    ("re:.*v_proj$", ["re:.*o_proj$"]),
    ("re:.*post_attention_layernorm$", ["re:.*mlp.experts.*.gate_proj$", "re:.*mlp.experts.*.up_proj$"]),
    ("re:.*up_proj$", ["re:.*down_proj$"]),
]


def default_mappings(names: Sequence[str]) -> List[Mapping]:
    """Llama / Qwen3-style defaults; the MoE table when the model has ``mlp.experts.<i>`` Linears (Qwen3-MoE family)."""
    moe = any(re.search(r"\.mlp\.experts\.\d+\.", n) for n in names)
    return [Mapping(s, list(b)) for s, b in (MOE_DEFAULT_MAPPINGS if moe else DEFAULT_MAPPINGS)]


def resolve_mappings(names: Sequence[str], spec: ModifierSpec) -> List[Tuple[str, List[str], str]]:
    """AWQ ``mappings`` -> concrete (smooth layer, balance layers, parent) triples (LLMC AWQModifier._set_resolved_mappings): every
    module matching ``smooth_layer`` is paired with the modules matching each ``balance_layers`` pattern that share the longest
    common dotted prefix with it (the same decoder layer / expert); the parent is the lowest common ancestor of the balance layers
    (a single balance layer is its own parent).  Mappings whose balance layers are all outside ``targets`` are dropped by the caller."""
    out = []
    for m in spec.mappings:
        # every pattern is matched against the module names ONCE (not once per smooth layer: 1e9 regex calls on a 62-layer,
        # 256-expert model otherwise)
        smooth_names = [n for n in names if match_name(n, m.smooth_layer)]
        cands_of = {pat: [n for n in names if match_name(n, pat)] for pat in m.balance_layers}
        for s_name in smooth_names:
            s_parts = s_name.split(".")
            balance: List[str] = []
            for pat in m.balance_layers:
                cands = [n for n in cands_of[pat] if n != s_name]
                if not cands:
                    continue

                def common(n):
                    k = 0
                    for a, b in zip(n.split("."), s_parts):
                        if a != b:
                            break
                        k += 1
                    return k

                best = max(common(n) for n in cands)
                balance += [n for n in cands if common(n) == best]
            if not balance:
                continue
            if len(balance) == 1:
                parent = balance[0]
            else:
                split = [b.split(".") for b in balance]
                k = 0
                while all(len(p) > k for p in split) and len({p[k] for p in split}) == 1:
                    k += 1
                parent = ".".join(split[0][:k])
            out.append((s_name, balance, parent))
    return out
