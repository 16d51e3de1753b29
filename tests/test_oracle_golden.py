"""CPU: pin the C oracle (oracle/ct_oracle.c) against the committed golden vectors that
tests/golden/make_golden.py recorded from live compressed_tensors 0.15.0.1."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.util import FORMATS, GOLDEN, assert_bits_equal, from_bits, geom_of, golden_files, load_golden

import os


@pytest.mark.parametrize("fname", golden_files())
def test_compress_matches_golden(fname):
    name, dtype, z = load_golden(fname)
    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    w = from_bits(z["w"], dtype)
    got = O.compress(w, fmt, geom_of(name), nb, sym)
    want_keys = {k[3:] for k in z if k.startswith("sd_")}
    assert set(got) == want_keys
    for k in want_keys:
        assert_bits_equal(got[k], z["sd_" + k], f"{name}:{k}")


@pytest.mark.parametrize("fname", golden_files())
def test_qparams_and_fake_quantize_match_golden(fname):
    name, dtype, z = load_golden(fname)
    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    w = from_bits(z["w"], dtype)
    geom = geom_of(name)
    mn, mx = O.minmax(w, geom)
    gs = torch.from_numpy(z["sd_weight_global_scale"]) if qtype == O.FP4 else None
    scale, zp = O.calculate_qparams(mn, mx, qtype, nb, sym, gs)
    scale_T = scale.to(dtype)  # update_weight_zp_scale copies into a Parameter of the weight dtype
    assert_bits_equal(scale_T.reshape(z["qp_scale"].shape), z["qp_scale"], f"{name}:scale")
    if "qp_zp" in z:
        assert_bits_equal(zp.reshape(z["qp_zp"].shape), z["qp_zp"], f"{name}:zp")
    fq = O.fake_quantize(w, scale_T, zp if qtype == O.INT else torch.zeros(1), geom, qtype, nb, gs)
    assert_bits_equal(fq, z["fq"], f"{name}:fake_quantize")


def test_kat_pack_and_fp4():
    z = np.load(os.path.join(GOLDEN, "kat.npz"))
    assert int(z["pack8_out"][0, 0]) == 33409  # CT:compressors/pack_quantized/helpers.py:28-38 docstring
    assert_bits_equal(O.pack_to_int32(torch.from_numpy(z["pack8_in"]), 8), z["pack8_out"], "pack8")
    v = torch.from_numpy(z["pack4_in"])
    assert_bits_equal(O.pack_to_int32(v, 4), z["pack4_out"], "pack4")
    assert_bits_equal(O.pack_to_int32(v, 4, 0), z["pack4_dim0_out"], "pack4 dim0")
    assert_bits_equal(O.unpack_from_int32(torch.from_numpy(z["pack4_out"]), 4, v.shape), z["pack4_in"], "unpack4")
    assert_bits_equal(O.unpack_from_int32(torch.from_numpy(z["pack4_dim0_out"]), 4, v.shape, 0), z["pack4_in"], "unpack4 d0")
    row = from_bits(z["fp4_row"], torch.bfloat16)
    q = O.quantize(row, torch.ones(1, 1, dtype=torch.bfloat16), torch.zeros(1), O.Geom(O.GROUP, 16), O.FP4, 4,
                   torch.ones(1))
    packed = (q[:, 0::2] | (q[:, 1::2] << 4)).to(torch.uint8)
    assert_bits_equal(packed, z["fp4_bytes"], "fp4 edge row")
    # hand-derived: [-0.1,0.1,-0.0,0.0,-0.25,0.25,-0.26,6.0] -> 0x08 0x00 0x08 0x79
    assert list(z["fp4_bytes"][0, :4]) == [0x08, 0x00, 0x08, 0x79]


@pytest.mark.parametrize("dn,dtype", [("bf16", torch.bfloat16), ("f16", torch.float16), ("f32", torch.float32)])
def test_kat_gparam(dn, dtype):
    z = np.load(os.path.join(GOLDEN, "kat.npz"))
    a = from_bits(z[f"gparam_in_{dn}"], dtype)
    want = z[f"gparam_out_{dn}"]
    got = np.array([O.generate_gparam(-float(x), float(x), dtype).item() for x in a.float()], dtype=np.float32)
    assert (got.view(np.int32) != want.view(np.int32)).sum() == 0


def test_e4m3_roundtrip_all_codes():
    L = O.lib()
    for c in range(256):
        if (c & 0x7F) == 0x7F:
            continue
        v = L.orc_e4m3_to_f32(c)
        assert L.orc_f32_to_e4m3(v) == c
        assert torch.tensor([v]).to(torch.float8_e4m3fn).view(torch.uint8).item() == c


def test_e4m3_rne_matches_torch():
    g = torch.Generator().manual_seed(5)
    x = torch.cat([torch.randn(20000, generator=g) * 100, torch.randn(20000, generator=g) * 0.01,
                   torch.linspace(-448, 448, 7001)]).clamp(-448, 448)
    want = x.to(torch.float8_e4m3fn).view(torch.uint8)
    L = O.lib()
    got = torch.tensor([L.orc_f32_to_e4m3(float(v)) for v in x], dtype=torch.uint8)
    assert (got != want).sum() == 0


def test_pack_unpack_roundtrip_ragged():
    g = torch.Generator().manual_seed(3)
    for R, C in [(1, 1), (7, 13), (9, 8), (16, 24)]:
        v = torch.randint(-8, 8, (R, C), generator=g, dtype=torch.int8)
        for dim in (0, 1):
            p = O.pack_to_int32(v, 4, dim)
            assert torch.equal(O.unpack_from_int32(p, 4, v.shape, dim), v)
