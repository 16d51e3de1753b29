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
# input-activation halves that need calibration (static); dynamic ones have no observer and nothing to calibrate
_PRESET_INPUTS: Dict[str, dict] = {
    "FP8": dict(num_bits=8, type="float", symmetric=True, strategy="tensor", dynamic=False, observer="memoryless_minmax"),
    "NVFP4": dict(num_bits=4, type="float", symmetric=True, strategy="tensor_group", group_size=16, dynamic="local", observer="static_minmax"),
}


class RecipeError(ValueError):
    pass


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
    a.observer = d.get("observer", "memoryless_minmax")
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
                                                      _args_from_dict(_PRESET_INPUTS[pname.upper()]) if pname.upper() in _PRESET_INPUTS else None,
                                                      preset=str(pname).upper()))
        else:
            key = str(scheme).upper()
            spec.config_groups.append(ConfigGroup(f"group_{len(spec.config_groups)}", targets, preset_args(key),
                                                  _args_from_dict(_PRESET_INPUTS[key]) if key in _PRESET_INPUTS else None, preset=key))
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
    """module name -> config group of this modifier (first matching group wins, like apply_quantization_config's ordered scan)."""
    out: Dict[str, ConfigGroup] = {}
    for name, mod in named_modules:
        for g in spec.config_groups:
            if is_match(name, mod, g.targets, spec.ignore):
                out[name] = g
                break
    return out


def resolve_mappings(names: Sequence[str], spec: ModifierSpec) -> List[Tuple[str, List[str], str]]:
    """AWQ ``mappings`` -> concrete (smooth layer, balance layers, parent) triples (LLMC AWQModifier._set_resolved_mappings): every
    module matching ``smooth_layer`` is paired with the modules matching each ``balance_layers`` pattern that share the longest
    common dotted prefix with it (the same decoder layer / expert); the parent is the lowest common ancestor of the balance layers
    (a single balance layer is its own parent).  Mappings whose balance layers are all outside ``targets`` are dropped by the caller."""
    out = []
    for m in spec.mappings:
        for s_name in names:
            if not match_name(s_name, m.smooth_layer):
                continue
            s_parts = s_name.split(".")
            balance: List[str] = []
            for pat in m.balance_layers:
                cands = [n for n in names if match_name(n, pat) and n != s_name]
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
