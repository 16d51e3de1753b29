"""CPU: mixed-precision recombination (REF:scripts/recombine_weights_MiniMax-M2.1.py:246-297 behaviour) on a toy pair of
checkpoints written with the safetensors package."""
import json
import os

import pytest
import torch

st = pytest.importorskip("safetensors.torch")


def _write(d, shards, extra_config=None):
    os.makedirs(d, exist_ok=True)
    wm = {}
    for f, tensors in shards.items():
        st.save_file(tensors, os.path.join(d, f), metadata={"format": "pt"})
        wm.update({k: f for k in tensors})
    with open(os.path.join(d, "model.safetensors.index.json"), "w") as fh:
        json.dump({"metadata": {}, "weight_map": wm}, fh)
    with open(os.path.join(d, "config.json"), "w") as fh:
        json.dump({"architectures": ["Toy"], **(extra_config or {})}, fh)


def test_recombine_minimax_rules(tmp_path):
    from safetensors import safe_open

    from quantizers_b200.recombine import recombine

    g = torch.Generator().manual_seed(0)
    f8 = lambda *s: torch.randn(*s, generator=g).to(torch.float8_e4m3fn)
    pre = "model.layers.0"
    base = {
        "model-00001-of-00130.safetensors": {
            f"{pre}.self_attn.q_proj.weight": f8(16, 32), f"{pre}.self_attn.q_proj.weight_scale_inv": torch.rand(1, 1, generator=g),
            f"{pre}.block_sparse_moe.experts.0.w1.weight": f8(8, 32), f"{pre}.block_sparse_moe.experts.0.w1.weight_scale_inv": torch.rand(1, 1, generator=g),
            f"{pre}.post_attention_layernorm.weight": torch.ones(32, dtype=torch.bfloat16),
        },
        "model-00002-of-00130.safetensors": {
            f"{pre}.block_sparse_moe.experts.1.w2.weight": f8(32, 8), f"{pre}.block_sparse_moe.experts.1.w2.weight_scale_inv": torch.rand(1, 1, generator=g),
            f"{pre}.block_sparse_moe.gate.weight": torch.randn(2, 32, generator=g).to(torch.bfloat16),
            "model.norm.weight": torch.ones(32, dtype=torch.bfloat16) * 3,
        },
    }
    over = {
        "model-00001-of-00002.safetensors": {
            f"{pre}.block_sparse_moe.experts.0.w1.weight_packed": torch.randint(-2**31, 2**31 - 1, (8, 4), generator=g, dtype=torch.int64).to(torch.int32),
            f"{pre}.block_sparse_moe.experts.0.w1.weight_scale": torch.rand(8, 1, generator=g).to(torch.bfloat16),
            f"{pre}.block_sparse_moe.experts.0.w1.weight_shape": torch.tensor([8, 32]),
            f"{pre}.post_attention_layernorm.weight": torch.full((32,), 0.5, dtype=torch.bfloat16),
        },
        "model-00002-of-00002.safetensors": {
            f"{pre}.block_sparse_moe.experts.1.w2.weight_packed": torch.randint(0, 1000, (32, 1), generator=g, dtype=torch.int64).to(torch.int32),
            f"{pre}.block_sparse_moe.experts.1.w2.weight_scale": torch.rand(32, 1, generator=g).to(torch.bfloat16),
            f"{pre}.block_sparse_moe.experts.1.w2.weight_shape": torch.tensor([32, 8]),
        },
    }
    b, o, out = str(tmp_path / "fp8"), str(tmp_path / "w4"), str(tmp_path / "out")
    _write(b, base)
    _write(o, over, {"quantization_config": {"ignore": ["lm_head"]}})
    dry = recombine(b, o, out, dry_run=True)
    assert not os.path.exists(out)
    stats = recombine(b, o, out)
    assert stats == dry
    assert (stats["pack_quantized_replaced"], stats["smoothing_layers_replaced"], stats["scale_inv_copied"], stats["scale_inv_skipped"]) == (2, 1, 1, 2)
    got = {}
    for f in sorted(os.listdir(out)):
        if f.endswith(".safetensors"):
            assert "00125" in f  # the reference's shard renaming
            with safe_open(os.path.join(out, f), framework="pt") as h:
                for k in h.keys():
                    got[k] = h.get_tensor(k)
    want = {}
    want[f"{pre}.self_attn.q_proj.weight"] = base["model-00001-of-00130.safetensors"][f"{pre}.self_attn.q_proj.weight"]
    want[f"{pre}.self_attn.q_proj.weight_scale"] = base["model-00001-of-00130.safetensors"][f"{pre}.self_attn.q_proj.weight_scale_inv"]
    for sh in over.values():
        want.update(sh)
    want[f"{pre}.block_sparse_moe.gate.weight"] = base["model-00002-of-00130.safetensors"][f"{pre}.block_sparse_moe.gate.weight"]
    want["model.norm.weight"] = base["model-00002-of-00130.safetensors"]["model.norm.weight"]
    assert set(got) == set(want)
    for k, v in want.items():
        assert got[k].dtype == v.dtype and torch.equal(got[k].view(torch.uint8), v.contiguous().view(torch.uint8)), k
    idx = json.load(open(os.path.join(out, "model.safetensors.index.json")))
    assert set(idx["weight_map"]) == set(want) and idx["metadata"]["total_size"] == stats["total_size"]
    cfg = json.load(open(os.path.join(out, "config.json")))
    qc = cfg["quantization_config"]
    assert cfg["architectures"] == ["Toy"] and qc["format"] == "mixed-precision" and qc["ignore"] == ["lm_head"]
    assert qc["config_groups"]["group_0"]["format"] == "float-quantized" and qc["config_groups"]["group_1"]["weights"]["group_size"] == 32
