"""model_free_ptq (REF:scripts/quant_GLM-4.7-Flash-FP8.py:11-24): safetensors container handling on the CPU, the streamed
quantization of a small sharded checkpoint on the GPU against the oracle."""
import json
import os

import pytest
import torch

from oracle import oracle as O
from tests.util import assert_bits_equal, synth_weight

safetensors = pytest.importorskip("safetensors")


def test_header_roundtrip_with_safetensors(tmp_path):
    """Files written with our header builder load with the safetensors package, and its files parse with our reader."""
    from safetensors import safe_open
    from safetensors.torch import save_file

    from quantizers_b200.model_free import build_header, read_header

    a = torch.arange(24, dtype=torch.float32).reshape(4, 6)
    b = (torch.randn(8, 16) * 3).to(torch.bfloat16)
    c = torch.tensor([5, 7], dtype=torch.int64)
    head, offs = build_header([("a", "F32", a.shape), ("m.weight", "BF16", b.shape), ("m.weight_shape", "I64", c.shape)], {"k": "v"})
    p = tmp_path / "x.safetensors"
    with open(p, "wb") as f:
        f.write(head)
        for t in (a, b, c):
            f.write(t.view(torch.uint8).numpy().tobytes() if t.dtype != torch.bfloat16 else t.view(torch.int16).numpy().tobytes())
    assert offs["a"][0] == len(head) and offs["m.weight"][0] == len(head) + a.numel() * 4
    with safe_open(str(p), framework="pt") as f:
        assert set(f.keys()) == {"a", "m.weight", "m.weight_shape"}
        assert torch.equal(f.get_tensor("a"), a) and torch.equal(f.get_tensor("m.weight"), b) and torch.equal(f.get_tensor("m.weight_shape"), c)
        assert f.metadata()["k"] == "v"
    q = tmp_path / "y.safetensors"
    save_file({"w": b, "z": a}, str(q), metadata={"format": "pt"})
    hdr, data0, meta = read_header(str(q))
    assert hdr["w"]["dtype"] == "BF16" and hdr["w"]["shape"] == [8, 16] and meta == {"format": "pt"}
    raw = open(q, "rb").read()
    o0, o1 = hdr["w"]["data_offsets"]
    assert raw[data0 + o0:data0 + o1] == b.view(torch.int16).numpy().tobytes()


def test_compressed_entries_layout():
    from quantizers_b200.model_free import compressed_entries
    from quantizers_b200.recipe import preset_args

    e = compressed_entries("m", 2560, 9728, "BF16", preset_args("W4A16_ASYM"))
    assert e == [("m.weight_packed", "I32", (2560, 1216)), ("m.weight_scale", "BF16", (2560, 76)), ("m.weight_zero_point", "I32", (320, 76)),
                 ("m.weight_shape", "I64", (2,))]
    e = compressed_entries("m", 200, 256, "BF16", preset_args("FP8_BLOCK"))
    assert e == [("m.weight", "F8_E4M3", (200, 256)), ("m.weight_scale", "BF16", (2, 2))]
    e = compressed_entries("m", 768, 2048, "BF16", preset_args("NVFP4"))
    assert e == [("m.weight_packed", "U8", (768, 1024)), ("m.weight_scale", "F8_E4M3", (768, 128)), ("m.weight_global_scale", "F32", (1,))]


def _checkpoint(d):
    from safetensors.torch import save_file

    t = {}
    for l in range(2):
        t[f"model.layers.{l}.self_attn.q_proj.weight"] = synth_weight(256, 384, torch.bfloat16, 10 + l)
        t[f"model.layers.{l}.self_attn.kv_a_proj_with_mqa.weight"] = synth_weight(64, 384, torch.bfloat16, 20 + l)
        t[f"model.layers.{l}.mlp.gate.weight"] = synth_weight(8, 384, torch.bfloat16, 30 + l)
        t[f"model.layers.{l}.mlp.experts.0.down_proj.weight"] = synth_weight(384, 200 + 56, torch.bfloat16, 40 + l)
        t[f"model.layers.{l}.mlp.experts.0.up_proj.weight"] = synth_weight(200, 384, torch.bfloat16, 50 + l)   # ragged rows for 128x128
        t[f"model.layers.{l}.input_layernorm.weight"] = torch.ones(384, dtype=torch.bfloat16) * (1 + l)
    t["model.embed_tokens.weight"] = synth_weight(100, 384, torch.bfloat16, 1)
    t["lm_head.weight"] = synth_weight(100, 384, torch.bfloat16, 2)
    names = sorted(t)
    shards = {"model-00001-of-00002.safetensors": names[: len(names) // 2], "model-00002-of-00002.safetensors": names[len(names) // 2:]}
    for f, ns in shards.items():
        save_file({n: t[n] for n in ns}, os.path.join(d, f), metadata={"format": "pt"})
    with open(os.path.join(d, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {"total_size": 0}, "weight_map": {n: f_ for f_, ns in shards.items() for n in ns}}, f)
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump({"architectures": ["Tiny"], "hidden_size": 384}, f)
    with open(os.path.join(d, "tokenizer.json"), "w") as f:
        f.write("{}")
    return t


IGNORE = ["lm_head", "re:.*mlp\\.gate$", "re:.*kv_a_proj_with_mqa$", "re:.*q_a_proj$", "model.embed_tokens"]


@pytest.mark.gpu
@pytest.mark.parametrize("scheme,fmt,geom,nb,sym", [("FP8_BLOCK", "float-quantized", O.Geom(O.BLOCK, 0, 128, 128), 8, True),
                                                   ("W4A16_ASYM", "pack-quantized", O.Geom(O.GROUP, 128), 4, False),
                                                   ("NVFP4", "nvfp4-pack-quantized", O.Geom(O.GROUP, 16), 4, True)])
def test_model_free_ptq_matches_oracle(tmp_path, scheme, fmt, geom, nb, sym):
    from safetensors import safe_open

    from quantizers_b200.model_free import model_free_ptq

    src, dst = tmp_path / "in", tmp_path / "out"
    os.makedirs(src)
    tensors = _checkpoint(str(src))
    stats = model_free_ptq(str(src), str(dst), scheme=scheme, ignore=IGNORE, max_workers=3, device="cuda:0")
    assert stats["files"] == 2 and stats["tensors_quantized"] == 6
    got = {}
    for f in sorted(os.listdir(dst)):
        if f.endswith(".safetensors"):
            with safe_open(os.path.join(dst, f), framework="pt") as h:
                for k in h.keys():
                    got[k] = h.get_tensor(k)
    index = json.load(open(dst / "model.safetensors.index.json"))
    assert set(index["weight_map"]) == set(got)
    for name, w in tensors.items():
        module = name[:-len(".weight")]
        quantized = w.ndim == 2 and not any(module == i or (i.startswith("re:") and __import__("re").match(i[3:], module)) for i in IGNORE)
        if not quantized:
            assert_bits_equal(got[name], w, f"pass-through {name}")
            continue
        want = O.compress(w, fmt, geom, nb, sym)
        for k, v in want.items():
            assert_bits_equal(got[f"{module}.{k}"].reshape(v.shape), v, f"{module}.{k}")
        assert name not in got or fmt == "float-quantized"
    cfg = json.load(open(dst / "config.json"))
    assert cfg["architectures"] == ["Tiny"] and cfg["quantization_config"]["format"] == fmt and cfg["quantization_config"]["ignore"] == IGNORE
    assert os.path.exists(dst / "tokenizer.json")


@pytest.mark.gpu
def test_model_free_ptq_rank_sharding(tmp_path):
    """Two ranks take alternate shards (no collective); together they produce the single-rank result."""
    from quantizers_b200.model_free import model_free_ptq

    src, one, two = tmp_path / "in", tmp_path / "one", tmp_path / "two"
    os.makedirs(src)
    _checkpoint(str(src))
    model_free_ptq(str(src), str(one), scheme="FP8_BLOCK", ignore=IGNORE, max_workers=2)
    s0 = model_free_ptq(str(src), str(two), scheme="FP8_BLOCK", ignore=IGNORE, max_workers=2, rank=0, world_size=2)
    s1 = model_free_ptq(str(src), str(two), scheme="FP8_BLOCK", ignore=IGNORE, max_workers=2, rank=1, world_size=2)
    assert s0["files"] == 1 and s1["files"] == 1
    for f in os.listdir(one):
        if f.endswith(".safetensors"):
            assert open(one / f, "rb").read() == open(two / f, "rb").read(), f
