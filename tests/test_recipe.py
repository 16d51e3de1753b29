"""CPU: recipe (YAML) parsing and target / mapping resolution -- host logic behind the reference's recipe schema
(REF:configs/recipes/*.yaml; the YAML below is typed here in that schema, not copied)."""
import os

import pytest
import torch

from quantizers_b200 import recipe as R

AWQ_LIST_FORM = """
quantization_scheme:
  type: W4A16
  targets: ["Linear"]
modifiers:
  - name: AWQModifier
    config_groups:
      group_0:
        targets: ["Linear"]
        weights: {num_bits: 4, type: int, symmetric: true, group_size: 32, strategy: group, dynamic: false, observer: memoryless_minmax}
    ignore: ["lm_head"]
    duo_scaling: true
"""

MIXED_STAGE_FORM = """
quant_stage:
  quant_modifiers:
    QuantizationModifier:
      targets: r"re:.*self_attn\\.(k|q|o|v)_proj.*"
      scheme: FP8_BLOCK
    AWQModifier:
      mlp_projections:
        group_0:
          targets: ["re:.*(down|gate|up)_proj.*"]
          weights: {num_bits: 4, type: int, symmetric: true, group_size: 32, strategy: group}
      ignore: ["lm_head"]
      duo_scaling: true
      mappings:
        - smooth_layer: re:.*post_attention_layernorm$
          balance_layers: ["re:.*gate_proj$", "re:.*up_proj$"]
        - smooth_layer: re:.*up_proj$
          balance_layers: ["re:.*down_proj$"]
"""

MOE_NVFP4 = """
default_stage:
  default_modifiers:
    QuantizationModifier:
      scheme: NVFP4
      targets:
        - "re:.*mlp\\\\.experts\\\\.\\\\d+\\\\.(down_proj|gate_proj|up_proj)$"
"""


class Block(torch.nn.Module):
    def __init__(self, h=32, i=64):
        super().__init__()
        self.input_layernorm = torch.nn.LayerNorm(h)
        self.post_attention_layernorm = torch.nn.LayerNorm(h)
        self.self_attn = torch.nn.Module()
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            setattr(self.self_attn, n, torch.nn.Linear(h, h, bias=False))
        self.mlp = torch.nn.Module()
        self.mlp.gate_proj = torch.nn.Linear(h, i, bias=False)
        self.mlp.up_proj = torch.nn.Linear(h, i, bias=False)
        self.mlp.down_proj = torch.nn.Linear(i, h, bias=False)


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.model = torch.nn.Module()
        self.model.layers = torch.nn.ModuleList([Block(), Block()])
        self.lm_head = torch.nn.Linear(32, 100, bias=False)


def test_list_form_and_ignore():
    rec = R.parse_recipe(AWQ_LIST_FORM)
    assert [m.kind for m in rec.modifiers] == ["AWQModifier"]
    spec = rec.modifier("AWQModifier")
    a = spec.config_groups[0].weights
    assert (a.num_bits, a.type, a.symmetric, a.strategy, a.group_size, a.format) == (4, "int", True, "group", 32, "pack-quantized")
    assert spec.ignore == ["lm_head"] and spec.duo_scaling is True and spec.n_grid == 20
    m = Tiny()
    lin = [(n, x) for n, x in m.named_modules() if isinstance(x, torch.nn.Linear)]
    got = R.resolve_targets(lin, spec)
    assert "lm_head" not in got and len(got) == 14


def test_stage_form_mixed_and_mappings():
    rec = R.parse_recipe(MIXED_STAGE_FORM)
    assert [m.kind for m in rec.modifiers] == ["QuantizationModifier", "AWQModifier"]
    q, a = rec.modifiers
    assert q.config_groups[0].preset == "FP8_BLOCK" and q.config_groups[0].weights.block_structure == [128, 128]
    assert q.config_groups[0].targets == ["re:.*self_attn\\.(k|q|o|v)_proj.*"]   # python-literal debris stripped
    assert a.config_groups[0].weights.group_size == 32 and len(a.mappings) == 2
    m = Tiny()
    lin = [(n, x) for n, x in m.named_modules() if isinstance(x, torch.nn.Linear)]
    assert sorted(R.resolve_targets(lin, q)) == sorted(n for n, _ in lin if "self_attn" in n)
    assert sorted(R.resolve_targets(lin, a)) == sorted(n for n, _ in lin if "mlp" in n)
    names = [n for n, _ in m.named_modules()]
    maps = R.resolve_mappings(names, a)
    assert ("model.layers.0.post_attention_layernorm", ["model.layers.0.mlp.gate_proj", "model.layers.0.mlp.up_proj"],
            "model.layers.0.mlp") in maps
    assert ("model.layers.1.mlp.up_proj", ["model.layers.1.mlp.down_proj"], "model.layers.1.mlp.down_proj") in maps
    assert len(maps) == 4


def test_moe_regex_targets_and_presets():
    spec = R.parse_recipe(MOE_NVFP4).modifier("QuantizationModifier")
    g = spec.config_groups[0]
    assert g.weights.format == "nvfp4-pack-quantized" and g.input_activations.observer == "static_minmax"
    lin = torch.nn.Linear(4, 4)
    named = [("model.layers.3.mlp.experts.17.gate_proj", lin), ("model.layers.3.mlp.gate", lin),
             ("model.layers.3.mlp.shared_expert.up_proj", lin), ("model.layers.3.self_attn.q_proj", lin)]
    assert list(R.resolve_targets(named, spec)) == ["model.layers.3.mlp.experts.17.gate_proj"]


def test_presets_match_compressed_tensors():
    ct = pytest.importorskip("compressed_tensors.quantization.quant_scheme")
    for name, kw in R._PRESET_WEIGHTS.items():
        w = ct.PRESET_SCHEMES[name].get("weights")
        if kw is None:
            assert w is None
            continue
        val = lambda v: getattr(v, "value", v)
        assert (w.num_bits, val(w.type), w.symmetric, val(w.strategy), w.group_size, w.block_structure) == (
            kw["num_bits"], kw["type"], kw["symmetric"], kw["strategy"], kw.get("group_size"), kw.get("block_structure")), name


def test_errors():
    with pytest.raises(ValueError):
        R.parse_recipe("foo: 1")
    with pytest.raises(ValueError):
        R.parse_recipe("modifiers:\n  - name: QuantizationModifier\n    scheme: W3A3\n")
    with pytest.raises(ValueError):
        R._args_from_dict(dict(num_bits=4, type="int", strategy="group"))
    with pytest.raises(ValueError):
        R._args_from_dict(dict(num_bits=4, type="int", strategy="group", group_size=128, actorder="group"))


_REF_RECIPES = "/root/reference/configs/recipes"


@pytest.mark.skipif(not os.path.isdir(_REF_RECIPES), reason="reference checkout not present (GPU box)")
def test_every_reference_recipe_parses():
    """The reference's own YAML recipes (REF:configs/recipes/*.yaml) go through the parser unchanged: modifier kinds, at least one
    config group with weight arguments each, AWQ mappings where the recipe defines them.  Read from the read-only reference checkout
    when it exists (this container); the restated recipes in the tests above cover the same schema on the GPU box."""
    from quantizers_b200 import recipe as R

    want = {
        "recipe_AR_W4A16G32.yaml": ["AutoRoundModifier"],
        "recipe_Dense_NVFP4.yaml": ["QuantizationModifier"],
        "recipe_Minimax-M2.1-AWQ-MixedPrec.yaml": ["AWQModifier"],
        "recipe_Minimax-M2.1-Experts-only-AWQ.yaml": ["AWQModifier"],
        "recipe_MoE_RTN_NVFP4.yaml": ["QuantizationModifier"],
        "recipe_awq_w4a16.yaml": ["AWQModifier"],
        "recipe_mixed_fp8_int4.yaml": ["QuantizationModifier", "AWQModifier"],
    }
    for fname, kinds in want.items():
        r = R.load_recipe(os.path.join(_REF_RECIPES, fname))
        assert [m.kind for m in r.modifiers] == kinds, fname
        for m in r.modifiers:
            assert m.config_groups, (fname, m.kind)
            for grp in m.config_groups:
                assert grp.weights is not None and grp.weights.num_bits in (4, 8), (fname, m.kind)
    awq = R.load_recipe(os.path.join(_REF_RECIPES, "recipe_Minimax-M2.1-Experts-only-AWQ.yaml")).modifiers[0]
    assert awq.mappings and len(awq.mappings) >= 2


def test_config_groups_serialise_like_compressed_tensors_presets():
    """quantization_config carries the FULL preset (weights + input activations, dynamic ones included) and the format
    compressed-tensors infers from both (FP8_BLOCK was written as W8A16 / INT+activations as pack-quantized)."""
    ct = pytest.importorskip("compressed_tensors")
    import json

    from compressed_tensors.compressors.format import infer_module_format
    from compressed_tensors.quantization.quant_scheme import preset_name_to_scheme
    import torch

    from quantizers_b200 import recipe as R

    for name in ("W8A16", "W4A16", "W4A16_ASYM", "W8A8", "INT8", "W4A8", "W4AFP8", "FP8", "FP8_DYNAMIC", "FP8_BLOCK", "NVFP4A16", "NVFP4"):
        scheme = preset_name_to_scheme(name, ["Linear"])
        want = json.loads(scheme.model_dump_json())
        grp = R.group_to_dict(R.ConfigGroup("g", ["Linear"], R.preset_args(name), R.preset_input_args(name), preset=name))
        for half in ("weights", "input_activations"):
            if want[half] is None:
                assert grp[half] is None, (name, half)
                continue
            for key in ("num_bits", "type", "symmetric", "group_size", "strategy", "block_structure", "dynamic", "observer"):
                assert grp[half][key] == want[half][key], (name, half, key, grp[half][key], want[half][key])
        assert grp["format"] == infer_module_format(torch.nn.Linear, scheme).value, name


def test_format_of_explicit_groups():
    from quantizers_b200 import recipe as R

    fp8_w = R._args_from_dict(dict(num_bits=8, type="float", strategy="group", group_size=32))
    assert R.infer_format(fp8_w, None) == "naive-quantized"          # FP8 weight-only (MiniMax mixed-precision AWQ recipe)
    int4 = R._args_from_dict(dict(num_bits=4, type="int", strategy="group", group_size=32))
    assert R.infer_format(int4, None) == "pack-quantized"
    assert R.infer_format(int4, R.preset_input_args("W4A8")) == "int-quantized"


def test_resolve_targets_priority_is_order_independent():
    """CT match_targets: name / regex matches beat class matches whatever the group order."""
    import torch

    from quantizers_b200 import recipe as R

    doc = """
quant_stage:
  quant_modifiers:
    QuantizationModifier:
      ignore: ["lm_head"]
      config_groups:
        everything:
          targets: ["Linear"]
          weights: {num_bits: 8, type: float, strategy: channel}
        mlp:
          targets: ["re:.*mlp.*"]
          weights: {num_bits: 4, type: int, strategy: group, group_size: 32}
"""
    spec = R.parse_recipe(doc).modifiers[0]
    mods = [("model.layers.0.self_attn.q_proj", torch.nn.Linear(4, 4)), ("model.layers.0.mlp.up_proj", torch.nn.Linear(4, 4)),
            ("lm_head", torch.nn.Linear(4, 4))]
    got = R.resolve_targets(mods, spec)
    assert got["model.layers.0.self_attn.q_proj"].name == "everything"
    assert got["model.layers.0.mlp.up_proj"].name == "mlp"
    assert "lm_head" not in got


def test_default_mappings_and_bucketed_resolution():
    from quantizers_b200 import recipe as R

    names = []
    for l in range(3):
        p = f"model.layers.{l}"
        names += [p, f"{p}.input_layernorm", f"{p}.self_attn", f"{p}.self_attn.q_proj", f"{p}.self_attn.k_proj", f"{p}.self_attn.v_proj",
                  f"{p}.self_attn.o_proj", f"{p}.post_attention_layernorm", f"{p}.mlp", f"{p}.mlp.gate_proj", f"{p}.mlp.up_proj",
                  f"{p}.mlp.down_proj"]
    spec = R.ModifierSpec(kind="AWQModifier", mappings=R.default_mappings(names))
    res = R.resolve_mappings(names, spec)
    assert len(res) == 3 * 4
    by_smooth = {s: (b, p) for s, b, p in res}
    b, p = by_smooth["model.layers.1.input_layernorm"]
    assert b == [f"model.layers.1.self_attn.{x}_proj" for x in "qkv"] and p == "model.layers.1.self_attn"
    b, p = by_smooth["model.layers.2.mlp.up_proj"]
    assert b == ["model.layers.2.mlp.down_proj"] and p == "model.layers.2.mlp.down_proj"
    moe = [n.replace("mlp.gate_proj", "mlp.experts.0.gate_proj") for n in names]
    assert R.default_mappings(moe)[2].balance_layers[0].startswith("re:.*mlp.experts")
