"""GPU: the oneshot-shaped driver (recipe -> modifiers -> compressed state dict) on a tiny decoder, against the oracle:
RTN outputs bit-exact per Linear, AWQ mappings resolved from the recipe with the same argmin / losses within 1e-3."""
import pytest
import torch

from oracle import llmc_restated as R
from oracle import oracle as O
from tests.util import assert_bits_equal

pytestmark = pytest.mark.gpu

H, I, HEADS = 128, 256, 4

RECIPE = """
quant_stage:
  quant_modifiers:
    QuantizationModifier:
      targets: "re:.*self_attn\\\\.(k|q|o|v)_proj.*"
      scheme: FP8_BLOCK
    AWQModifier:
      mlp_projections:
        group_0:
          targets: ["re:.*(down|gate|up)_proj.*"]
          weights: {num_bits: 4, type: int, symmetric: true, group_size: 32, strategy: group, observer: memoryless_minmax}
      ignore: ["lm_head"]
      duo_scaling: true
      mappings:
        - smooth_layer: re:.*post_attention_layernorm$
          balance_layers: ["re:.*gate_proj$", "re:.*up_proj$"]
        - smooth_layer: re:.*up_proj$
          balance_layers: ["re:.*down_proj$"]
"""

NVFP4_RECIPE = """
default_stage:
  default_modifiers:
    QuantizationModifier:
      scheme: NVFP4
      targets: ["Linear"]
      ignore: ["lm_head"]
"""


class Attn(torch.nn.Module):
    def __init__(self):
        super().__init__()
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            setattr(self, n, torch.nn.Linear(H, H, bias=False))

    def forward(self, x):
        B, S, _ = x.shape
        q, k, v = (p(x).view(B, S, HEADS, H // HEADS).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.o_proj(o.transpose(1, 2).reshape(B, S, H))


class MLP(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.gate_proj = torch.nn.Linear(H, I, bias=False)
        self.up_proj = torch.nn.Linear(H, I, bias=False)
        self.down_proj = torch.nn.Linear(I, H, bias=False)

    def forward(self, x):
        return self.down_proj(torch.nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))


class Block(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.input_layernorm = torch.nn.LayerNorm(H)
        self.post_attention_layernorm = torch.nn.LayerNorm(H)
        self.self_attn, self.mlp = Attn(), MLP()

    def forward(self, x):
        x = x + self.self_attn(self.input_layernorm(x))
        return x + self.mlp(self.post_attention_layernorm(x))


class TinyLM(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.model = torch.nn.Module()
        self.model.layers = torch.nn.ModuleList([Block(), Block()])
        self.lm_head = torch.nn.Linear(H, 64, bias=False)

    def forward(self, x):
        for b in self.model.layers:
            x = b(x)
        return self.lm_head(x)


def _model(seed=0):
    torch.manual_seed(seed)
    m = TinyLM()
    for p in m.parameters():
        if p.ndim == 2:
            p.data.normal_(0, 0.05)
            p.data[:, 3] *= 8
    return m.to(torch.bfloat16).cuda()


def _batches(n=4, S=96, seed=1):
    g = torch.Generator().manual_seed(seed)
    spread = 1 + 3 * torch.rand(H, generator=g)
    return [(torch.randn(1, S, H, generator=g) * spread).to(torch.bfloat16).cuda() for _ in range(n)]


def test_oneshot_rtn_nvfp4_fused_siblings():
    """QuantizationModifier NVFP4 on every Linear: q/k/v and gate/up share min(global_scale); every tensor equals the oracle."""
    from quantizers_b200.oneshot import oneshot

    m = _model(3)
    sd, cfg = oneshot(m, NVFP4_RECIPE)
    assert cfg["format"] == "nvfp4-pack-quantized" and cfg["ignore"] == ["lm_head"]
    assert not any(k.startswith("lm_head") for k in sd)
    mods = dict(m.named_modules())
    geom = O.Geom(O.GROUP, 16)
    for layer in range(2):
        for parent, sibs in ((f"model.layers.{layer}.self_attn", ("q_proj", "k_proj", "v_proj")), (f"model.layers.{layer}.mlp", ("gate_proj", "up_proj")),
                             (f"model.layers.{layer}.self_attn", ("o_proj",)), (f"model.layers.{layer}.mlp", ("down_proj",))):
            ws = [mods[f"{parent}.{s}"].weight.detach().cpu() for s in sibs]
            gs = min(float(O.generate_gparam(float(w.float().min()), float(w.float().max()), torch.bfloat16)) for w in ws)
            for s, w in zip(sibs, ws):
                want = O.compress(w, "nvfp4-pack-quantized", geom, 4, True, torch.tensor([gs]))
                for k, v in want.items():
                    assert_bits_equal(sd[f"{parent}.{s}.{k}"].reshape(v.shape), v, f"{parent}.{s}.{k}")


def test_oneshot_mixed_fp8_block_and_awq_int4():
    """The reference's mixed recipe shape (REF:configs/recipes/recipe_mixed_fp8_int4.yaml): FP8_BLOCK RTN on attention, AWQ INT4 g32
    on the MLP with explicit mappings.  Mapping resolution, search results and the final packed tensors are checked."""
    from quantizers_b200 import recipe as RC
    from quantizers_b200.oneshot import _Capture, awq_model, quantize_model

    m = _model(5)
    batches = _batches()
    rec = RC.parse_recipe(RECIPE)
    # what the search sees: inputs of the balance layers on the un-smoothed model
    names = [f"model.layers.{l}.mlp.{p}" for l in range(2) for p in ("gate_proj", "down_proj")]
    cap = _Capture(m, names, [])
    with torch.no_grad():
        for b in batches:
            m(b)
    cap.close()
    x_cpu = {n: [t.cpu() for t in v] for n, v in cap.inputs.items()}
    w0 = {n: p.detach().cpu().clone() for n, p in m.named_parameters()}

    sd_a, cfg_a, results = awq_model(m, rec, batches)
    sd_q, cfg_q = quantize_model(m, rec)
    assert cfg_q["format"] == "float-quantized" and cfg_a["format"] == "pack-quantized"
    assert len(results) == 4
    geom = O.Geom(O.GROUP, 32)
    for l in range(2):
        pre = f"model.layers.{l}.mlp"
        s_ref, r_ref, l_ref = R.compute_best_scale(x_cpu[f"{pre}.gate_proj"], [w0[f"{pre}.gate_proj.weight"], w0[f"{pre}.up_proj.weight"]],
                                                   R.mlp_parent(w0[f"{pre}.down_proj.weight"]), geom, O.INT, 4, True)
        s, r, losses = results[f"model.layers.{l}.post_attention_layernorm -> {pre}.gate_proj,{pre}.up_proj"]
        assert r == r_ref and max(abs(a - b) / b for a, b in zip(losses, l_ref)) < 1e-3
        assert torch.allclose(s, s_ref, rtol=1e-5)
        # up -> down runs on the smoothed up_proj's output; check the fold-in of the first mapping and the final RTN tensors
        new_w, new_ln = R.smooth([w0[f"{pre}.gate_proj.weight"]], w0[f"model.layers.{l}.post_attention_layernorm.weight"], s)
        s2, _, _ = results[f"{pre}.up_proj -> {pre}.down_proj"]
        assert_bits_equal(dict(m.named_parameters())[f"{pre}.gate_proj.weight"].detach(), new_w[0], "gate smoothed")
        assert_bits_equal(dict(m.named_parameters())[f"model.layers.{l}.post_attention_layernorm.weight"].detach(), new_ln, "ln smoothed")
        for p in ("gate_proj", "up_proj", "down_proj"):
            w = dict(m.named_parameters())[f"{pre}.{p}.weight"].detach().cpu()
            want = O.compress(w, "pack-quantized", geom, 4, True)
            for k, v in want.items():
                assert_bits_equal(sd_a[f"{pre}.{p}.{k}"].reshape(v.shape), v, f"{pre}.{p}.{k}")
        for p in ("q_proj", "k_proj", "v_proj", "o_proj"):
            w = w0[f"model.layers.{l}.self_attn.{p}.weight"]
            want = O.compress(w, "float-quantized", O.Geom(O.BLOCK, 0, 128, 128), 8, True)
            for k, v in want.items():
                assert_bits_equal(sd_q[f"model.layers.{l}.self_attn.{p}.{k}"].reshape(v.shape), v, f"attn {p}.{k}")


def test_oneshot_refuses_cpu_weights():
    from quantizers_b200.oneshot import oneshot

    m = TinyLM().to(torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        oneshot(m, NVFP4_RECIPE)


# ----------------------------------------------------------------------------- MoE recipe (config 5 shape) through awq_model
MOE_RECIPE = """
default_stage:
  default_modifiers:
    AWQModifier:
      config_groups:
        mlp_experts_projections:
          targets: ["re:.*block_sparse_moe\\\\.experts\\\\.\\\\d+\\\\.(w1|w2|w3)$"]
          weights: {num_bits: 4, type: int, symmetric: true, group_size: 32, strategy: group, dynamic: false, observer: memoryless_minmax}
      mappings:
        - smooth_layer: re:.*post_attention_layernorm$
          balance_layers: ["re:.*w1$", "re:.*w3$"]
        - smooth_layer: re:.*w3$
          balance_layers: ["re:.*w2$"]
      duo_scaling: true
"""


class Expert(torch.nn.Module):
    def __init__(self, h, i):
        super().__init__()
        self.w1, self.w3, self.w2 = torch.nn.Linear(h, i, bias=False), torch.nn.Linear(h, i, bias=False), torch.nn.Linear(i, h, bias=False)

    def forward(self, x):
        return self.w2(torch.nn.functional.silu(self.w1(x)) * self.w3(x))


class SparseMoE(torch.nn.Module):
    """Routed block in "calibrate all experts" form (what llmcompressor swaps in, REF:scripts/do_oneshot.py:186): every expert
    runs on ALL tokens, only the routed rows reach the output (index_add_ per expert, ascending)."""

    def __init__(self, h=128, i=256, e=4, k=2):
        super().__init__()
        self.gate = torch.nn.Linear(h, e, bias=False)
        self.experts = torch.nn.ModuleList([Expert(h, i) for _ in range(e)])
        self.k = k

    def forward(self, x):
        B, S, H = x.shape
        xf = x.reshape(-1, H)
        p = torch.softmax(self.gate(xf).float(), dim=-1)
        w, idx = torch.topk(p, self.k, dim=-1)
        w = (w / w.sum(-1, keepdim=True)).to(x.dtype)
        out = torch.zeros_like(xf)
        for e, ex in enumerate(self.experts):
            y = ex(xf)
            tok, slot = torch.where(idx == e)
            if tok.numel():
                out.index_add_(0, tok, y[tok] * w[tok, slot, None])
        return out.view(B, S, H)


class MoELayer(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.post_attention_layernorm = torch.nn.LayerNorm(128)
        self.block_sparse_moe = SparseMoE()

    def forward(self, x):
        return x + self.block_sparse_moe(self.post_attention_layernorm(x))


class TinyMoELM(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.model = torch.nn.Module()
        self.model.layers = torch.nn.ModuleList([MoELayer()])

    def forward(self, x):
        return self.model.layers[0](x)


def test_awq_moe_recipe_through_oneshot():
    """MiniMax-style experts-only AWQ recipe (REF:configs/recipes/recipe_Minimax-M2.1-Experts-only-AWQ.yaml shape): the layer-wide
    mapping resolves to all experts' w1 + w3 with the sparse-MoE block as parent, the per-expert mappings to w3 -> w2; results
    against the restated search replaying the same block on the CPU."""
    import copy

    from torch.func import functional_call

    from quantizers_b200 import recipe as RC
    from quantizers_b200.oneshot import _Capture, awq_model

    torch.manual_seed(11)
    m = TinyMoELM()
    for p in m.parameters():
        if p.ndim == 2:
            p.data.normal_(0, 0.05)
    m.model.layers[0].block_sparse_moe.gate.weight.data.normal_(0, 1.0)
    m = m.to(torch.bfloat16).cuda()
    cpu = copy.deepcopy(m).cpu()
    batches = _batches(n=3, S=64, seed=5)
    pre = "model.layers.0.block_sparse_moe"
    names = [f"{pre}.experts.0.w1"] + [f"{pre}.experts.{e}.w2" for e in range(4)]
    cap = _Capture(m, names, [pre])
    with torch.no_grad():
        for b in batches:
            m(b)
    cap.close()
    w0 = {n: p.detach().cpu().clone() for n, p in m.named_parameters()}
    sd, cfg, results = awq_model(m, RC.parse_recipe(MOE_RECIPE), batches)
    assert len(results) == 1 + 4 and cfg["format"] == "pack-quantized"
    # layer-wide mapping: balance = w1 of every expert, then w3 of every expert; parent = the MoE block
    balance = [f"{pre}.experts.{e}.w1" for e in range(4)] + [f"{pre}.experts.{e}.w3" for e in range(4)]
    key = "model.layers.0.post_attention_layernorm -> " + ",".join(balance)
    assert key in results
    block = cpu.model.layers[0].block_sparse_moe
    calls = [a[0].cpu() for a, _ in cap.parent_args[pre]]
    rel = [b[len(pre) + 1:] + ".weight" for b in balance]

    def parent(weights, _x):
        outs = [functional_call(block, dict(zip(rel, weights)), (c,)) for c in calls]
        return torch.cat([o.reshape(-1, o.shape[-1]) for o in outs])

    x_cpu = torch.cat([t.cpu() for t in cap.inputs[f"{pre}.experts.0.w1"]])
    s_ref, r_ref, l_ref = R.compute_best_scale([x_cpu], [w0[b + ".weight"] for b in balance], parent, O.Geom(O.GROUP, 32), O.INT, 4, True)
    s, r, losses = results[key]
    assert r == r_ref and max(abs(a - b) / b for a, b in zip(losses, l_ref)) < 1e-3
    assert torch.allclose(s, s_ref, rtol=1e-5)
    # every target got packed tensors
    for e in range(4):
        for wn in ("w1", "w2", "w3"):
            assert f"{pre}.experts.{e}.{wn}.weight_packed" in sd and f"{pre}.experts.{e}.{wn}.weight_scale" in sd
    assert not any(k.startswith(f"{pre}.gate") for k in sd)


def test_oneshot_nvfp4_input_global_scales():
    """NVFP4 scheme with calibration data: every target also gets ``input_global_scale`` from the static_minmax observer over
    all batches (LLMC calibrate_activations; CT:quantization/quant_scheme.py:170-180)."""
    from quantizers_b200.oneshot import _Capture, oneshot

    m = _model(7)
    batches = _batches(n=3, S=48, seed=9)
    names = [n for n, mod in m.named_modules() if isinstance(mod, torch.nn.Linear) and n != "lm_head"]
    cap = _Capture(m, names, [])
    with torch.no_grad():
        for b in batches:
            m(b)
    cap.close()
    sd, cfg = oneshot(m, NVFP4_RECIPE, dataset=batches)
    for n in names:
        want = R.activation_global_scale([t.cpu() for t in cap.inputs[n]])
        got = sd[f"{n}.input_global_scale"]
        assert got.dtype == torch.float32 and got.numel() == 1 and got.item() == want.item(), n
    assert "lm_head.input_global_scale" not in sd


def test_save_compressed_roundtrip(tmp_path):
    """oneshot -> save_compressed -> files load with the safetensors package: compressed tensors in place of the quantized
    weights, untouched parameters as they were, sharded index and quantization_config present."""
    import json
    import os

    st = pytest.importorskip("safetensors")
    from quantizers_b200.oneshot import oneshot, save_compressed

    m = _model(13)
    sd, cfg = oneshot(m, NVFP4_RECIPE)
    info = save_compressed(m, sd, cfg, str(tmp_path), config={"architectures": ["TinyLM"]}, max_shard_bytes=200_000)
    assert info["files"] > 1
    got = {}
    for f in os.listdir(tmp_path):
        if f.endswith(".safetensors"):
            with st.safe_open(os.path.join(tmp_path, f), framework="pt") as h:
                for k in h.keys():
                    got[k] = h.get_tensor(k)
    idx = json.load(open(tmp_path / "model.safetensors.index.json"))
    assert set(idx["weight_map"]) == set(got)
    for k, v in sd.items():
        a, b = got[k], v.cpu()
        assert a.dtype == b.dtype and torch.equal(a.view(torch.uint8), b.contiguous().view(torch.uint8)), k
    assert "model.layers.0.self_attn.q_proj.weight" not in got and "model.layers.0.self_attn.q_proj.weight_packed" in got
    assert torch.equal(got["lm_head.weight"], m.lm_head.weight.detach().cpu())
    assert torch.equal(got["model.layers.1.input_layernorm.weight"], m.model.layers[1].input_layernorm.weight.detach().cpu())
    c = json.load(open(tmp_path / "config.json"))
    assert c["architectures"] == ["TinyLM"] and c["quantization_config"]["format"] == "nvfp4-pack-quantized"


# ----------------------------------------------------------------------------- CompressedLinear: run from the compressed tensors
MIXED_RTN_RECIPE = """
default_stage:
  default_modifiers:
    QuantizationModifier:
      config_groups:
        attn:
          targets: ["re:.*self_attn\\\\.(q|k|v|o)_proj$"]
          weights: {num_bits: 8, type: float, symmetric: true, strategy: block, block_structure: [128, 128], observer: memoryless_minmax}
        mlp:
          targets: ["re:.*mlp\\\\.(gate|up|down)_proj$"]
          weights: {num_bits: 4, type: int, symmetric: false, group_size: 128, strategy: group, observer: memoryless_minmax}
      ignore: ["lm_head"]
"""


def _oracle_fake_quantized(w, geom, qtype, bits, sym, gs=None):
    mn, mx = O.minmax(w, geom)
    scale, zp = O.calculate_qparams(mn, mx, qtype, bits, sym, gs)
    if gs is not None:
        scale = scale.to(torch.float8_e4m3fn)
    zp = None if sym and qtype == O.INT else zp
    fq = O.fake_quantize(w, scale, zp, geom, qtype, bits, gs)
    if qtype != O.INT:
        return fq
    # integer codes cannot carry the sign of a zero: fake_quantize keeps round(-0.3) = -0.0 through (q - 0) * s, the decode of
    # the stored code 0 gives +0.0 -- equal as numbers, different as bits (live CT behaves the same way)
    dq = O.dequantize(O.quantize(w, scale, zp, geom, qtype, bits), scale, zp, geom, qtype)
    assert torch.equal(dq, fq)
    return dq


@pytest.mark.parametrize("recipe", ["mixed", "nvfp4"])
def test_compressed_linear_decodes_to_fake_quantized_weight(recipe):
    """oneshot -> apply_compressed: every swapped Linear decodes (per forward call) to exactly fake_quantize(original weight), so
    the model output equals the output of the dense model carrying the oracle's fake-quantized weights."""
    import copy

    from quantizers_b200.compressed_linear import CompressedLinear, apply_compressed
    from quantizers_b200.oneshot import oneshot

    m = _model(9)
    dense = copy.deepcopy(m)
    text = MIXED_RTN_RECIPE if recipe == "mixed" else NVFP4_RECIPE
    sd, _ = oneshot(m, text)
    swapped = apply_compressed(m, text, sd)
    assert len(swapped) == 14 and "lm_head" not in swapped
    mods, dmods = dict(m.named_modules()), dict(dense.named_modules())
    for name in swapped:
        cl = mods[name]
        assert isinstance(cl, CompressedLinear)
        w = dmods[name].weight.detach().cpu()
        if recipe == "nvfp4":
            gs = sd[f"{name}.weight_global_scale"].reshape(1).cpu()
            want = _oracle_fake_quantized(w, O.Geom(O.GROUP, 16), O.FP4, 4, True, gs)
        elif "self_attn" in name:
            want = _oracle_fake_quantized(w, O.Geom(O.BLOCK, 0, 128, 128), O.FP8, 8, True)
        else:
            want = _oracle_fake_quantized(w, O.Geom(O.GROUP, 128), O.INT, 4, False)
        assert_bits_equal(cl.decompressed_weight(), want, name)
        dmods[name].weight.data.copy_(want.cuda())
    x = _batches(n=1, S=32, seed=3)[0]
    with torch.no_grad():
        assert_bits_equal(m(x), dense(x), "forward from compressed tensors")
    # the dense weights are gone: only compressed buffers remain on the swapped modules
    kept = sum(b.numel() * b.element_size() for n in swapped for b in mods[n].buffers())
    full = sum(dmods[n].weight.numel() * 2 for n in swapped)
    assert kept < 0.6 * full


def test_compressed_linear_rejects_bad_entries():
    from quantizers_b200.compressed_linear import CompressedLinear
    from quantizers_b200.scheduler import PRESETS

    with pytest.raises(ValueError):
        CompressedLinear({"weight_scale": torch.zeros(1)}, PRESETS["W4A16"], 8, 8)
    with pytest.raises(ValueError):
        CompressedLinear({"weight_packed": torch.zeros(1), "weight_scale": torch.zeros(1), "bogus": torch.zeros(1)}, PRESETS["W4A16"], 8, 8)


# ----------------------------------------------------------------------------- moe_calibrate_all_experts through oneshot
HF_MOE_RECIPE = """
default_stage:
  default_modifiers:
    AWQModifier:
      config_groups:
        experts:
          targets: ["re:.*mlp\\\\.experts\\\\.\\\\d+\\\\.(gate|up|down)_proj$"]
          weights: {num_bits: 4, type: int, symmetric: true, group_size: 32, strategy: group, dynamic: false, observer: memoryless_minmax}
      mappings:
        - smooth_layer: re:.*up_proj$
          balance_layers: ["re:.*down_proj$"]
      duo_scaling: true
"""


def test_oneshot_moe_calibrate_all_experts_on_fused_experts():
    """transformers' Qwen3-MoE block stores its experts as fused 3-D parameters: invisible to Linear targets
    (REF:docs/quantization_tips_and_tricks.md:77-80).  With ``moe_calibrate_all_experts=True`` (REF:scripts/do_oneshot.py:186) the
    block is linearized, every expert's up -> down mapping is searched on ALL calibration tokens, and the result matches the
    restated search on those inputs."""
    qm = pytest.importorskip("transformers.models.qwen3_moe.modeling_qwen3_moe")
    from transformers import Qwen3MoeConfig

    from quantizers_b200 import recipe as RC
    from quantizers_b200.oneshot import awq_model
    from quantizers_b200.moe_calibration import moe_calibrate_all_experts

    Hh, Ii, E = 128, 128, 4
    cfg = Qwen3MoeConfig(hidden_size=Hh, moe_intermediate_size=Ii, num_experts=E, num_experts_per_tok=2, norm_topk_prob=True,
                         num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=2, vocab_size=32)

    class Layer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.post_attention_layernorm = torch.nn.LayerNorm(Hh)
            self.mlp = qm.Qwen3MoeSparseMoeBlock(cfg)

        def forward(self, x):
            y = self.mlp(self.post_attention_layernorm(x))
            return x + (y[0] if isinstance(y, tuple) else y)

    class LM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.Module()
            self.model.layers = torch.nn.ModuleList([Layer()])

        def forward(self, x):
            return self.model.layers[0](x)

    torch.manual_seed(21)
    m = LM()
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim >= 2:
                p.normal_(0, 0.06)
        m.model.layers[0].mlp.gate.weight.normal_(0, 1.0)
    m = m.to(torch.bfloat16).cuda()
    g = torch.Generator().manual_seed(2)
    batches = [(torch.randn(1, 80, Hh, generator=g)).to(torch.bfloat16).cuda() for _ in range(3)]
    rec = RC.parse_recipe(HF_MOE_RECIPE)
    sd0, _, res0 = awq_model(m, rec, batches)            # fused experts: nothing for the recipe to bind to
    assert not sd0 and not res0
    with torch.no_grad():
        before = m(batches[0]).clone()
    with moe_calibrate_all_experts(m) as names:
        assert names == ["model.layers.0.mlp"]
        with torch.no_grad():
            assert torch.allclose(m(batches[0]).float(), before.float(), rtol=2e-2, atol=2e-2)
        pre = "model.layers.0.mlp.experts"
        seen = {e: [] for e in range(E)}
        hooks = [m.get_submodule(f"{pre}.{e}.down_proj").register_forward_pre_hook(lambda mod, a, e=e: seen[e].append(a[0].detach().cpu()))
                 for e in range(E)]
        with torch.no_grad():
            for b in batches:
                m(b)
        for h in hooks:
            h.remove()
        w0 = {e: m.get_submodule(f"{pre}.{e}.down_proj").weight.detach().cpu().clone() for e in range(E)}
        sd, qcfg, results = awq_model(m, rec, batches)
    assert len(results) == E
    for e in range(E):
        x = torch.cat(seen[e])
        assert x.shape[0] == 3 * 80                       # every expert saw every token
        s_ref, r_ref, l_ref = R.compute_best_scale([x], [w0[e]], R.linear_parent, O.Geom(O.GROUP, 32), O.INT, 4, True)
        s, r, losses = results[f"{pre}.{e}.up_proj -> {pre}.{e}.down_proj"]
        assert r == r_ref and max(abs(a - b) / b for a, b in zip(losses, l_ref)) < 1e-3
        assert torch.allclose(s, s_ref, rtol=1e-5)
        for p in ("gate_proj", "up_proj", "down_proj"):
            assert f"{pre}.{e}.{p}.weight_packed" in sd


# REF:configs/recipes/recipe_awq_w4a16.yaml restated (same keys, same layout): NO `mappings:` -- llmcompressor's model-family
# defaults apply (input_layernorm -> q/k/v, v -> o [dropped under GQA], post_attention_layernorm -> gate/up, up -> down)
AWQ_DEFAULTS_RECIPE = """
quantization_scheme:
  type: W4A16
  targets: ["Linear"]

modifiers:
  - name: AWQModifier
    config_groups:
      group_0:
        targets: ["Linear"]
        weights:
          num_bits: 4
          type: int
          symmetric: true
          group_size: 32
          strategy: group
          dynamic: false
          observer: memoryless_minmax
    ignore:
      - "lm_head"
    duo_scaling: true
"""


class GQAAttn(torch.nn.Module):
    """Grouped-query attention: k / v project to HEADS_KV * d < H, so v_proj.out_features != o_proj.in_features."""

    def __init__(self, kv_heads):
        super().__init__()
        d = H // HEADS
        self.kv_heads = kv_heads
        self.q_proj = torch.nn.Linear(H, H, bias=False)
        self.k_proj = torch.nn.Linear(H, kv_heads * d, bias=False)
        self.v_proj = torch.nn.Linear(H, kv_heads * d, bias=False)
        self.o_proj = torch.nn.Linear(H, H, bias=False)

    def forward(self, x):
        B, S, _ = x.shape
        d = H // HEADS
        q = self.q_proj(x).view(B, S, HEADS, d).transpose(1, 2)
        k = self.k_proj(x).view(B, S, self.kv_heads, d).transpose(1, 2)
        v = self.v_proj(x).view(B, S, self.kv_heads, d).transpose(1, 2)
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=self.kv_heads != HEADS)
        return self.o_proj(o.transpose(1, 2).reshape(B, S, H))


@pytest.mark.parametrize("kv_heads", [HEADS, HEADS // 2])
def test_awq_recipe_without_mappings_uses_family_defaults(kv_heads):
    """the reference's main AWQ recipe has no `mappings:`; oneshot() must run it end to end.  MHA keeps v -> o, GQA drops
    it (shapes cannot be folded).  Every search is checked against the restated oracle on the weights it actually saw."""
    from quantizers_b200 import recipe as RC
    from quantizers_b200.oneshot import oneshot, awq_model

    m = _model(9)
    for blk in m.model.layers:
        blk.self_attn = GQAAttn(kv_heads).to(torch.bfloat16).cuda()
        for p in blk.self_attn.parameters():
            p.data.normal_(0, 0.05)
    batches = _batches()
    rec = RC.parse_recipe(AWQ_DEFAULTS_RECIPE)
    assert [mod.kind for mod in rec.modifiers] == ["AWQModifier"] and not rec.modifiers[0].mappings
    sd, cfg, results = awq_model(m, rec, batches)
    per_layer = 4 if kv_heads == HEADS else 3
    assert len(results) == 2 * per_layer, sorted(results)
    has_vo = any("v_proj -> " in k for k in results)
    assert has_vo == (kv_heads == HEADS)
    for k, (s, r, losses) in results.items():
        assert 0.0 <= r < 1.0 and all(l == l and l >= 0 for l in losses), k
    # all seven projections of both layers quantized, lm_head ignored, format as compressed-tensors infers it
    assert cfg["format"] == "pack-quantized" and cfg["ignore"] == ["lm_head"]
    for l in range(2):
        for name in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"):
            w = dict(m.named_parameters())[f"model.layers.{l}.{name}.weight"].detach().cpu()
            want = O.compress(w, "pack-quantized", O.Geom(O.GROUP, 32), 4, True)
            for key, v in want.items():
                assert_bits_equal(sd[f"model.layers.{l}.{name}.{key}"].reshape(v.shape), v, f"{name}.{key}")
    assert not any(k.startswith("lm_head.weight_packed") for k in sd)
    # and through the oneshot() entry point itself
    m2 = _model(9)
    sd2, cfg2 = oneshot(m2, AWQ_DEFAULTS_RECIPE, dataset=batches)
    assert cfg2["config_groups"]["group_0"]["weights"]["group_size"] == 32
