"""CPU: the arithmetic header the kernels inline (quantizers_b200/csrc/qmath.cuh), compiled for the host, must
agree bit-for-bit with the oracle.  This is what lets kernel math be validated without a GPU."""
import ctypes
import os

import pytest
import torch

from oracle import oracle as O
from tests.util import FORMATS, assert_bits_equal, geom_of, synth_weight

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "quantizers_b200", "lib", "libb200q_hostmath.so")


@pytest.fixture(scope="module")
def hm():
    if not os.path.exists(SO):
        from quantizers_b200.build import build

        build()
    lib = ctypes.CDLL(SO)
    lib.hm_gparam.restype = ctypes.c_float
    lib.hm_gparam.argtypes = [ctypes.c_float, ctypes.c_int]
    lib.hm_e4m3_encode.restype = ctypes.c_uint8
    lib.hm_e4m3_encode.argtypes = [ctypes.c_float]
    lib.hm_e4m3_decode.restype = ctypes.c_float
    lib.hm_e4m3_decode.argtypes = [ctypes.c_uint8]
    return lib


GROUP_FORMATS = [n for n, f in FORMATS.items() if f[4] == O.GROUP]


@pytest.mark.parametrize("name", GROUP_FORMATS)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_group_compress_and_fq(hm, name, dtype):
    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    R, C = 24, 512
    w = synth_weight(R, C, dtype, 42)
    G = C // g
    codes = torch.empty((R, C), dtype=torch.uint8)
    scale = torch.empty((R, G), dtype=torch.uint8 if qtype == O.FP4 else dtype)
    zp = torch.zeros((R, G), dtype=torch.int8)
    gs_out = torch.zeros(1, dtype=torch.float32)
    rc = hm.hm_compress_group(ctypes.c_void_p(w.data_ptr()), O._DT[dtype], ctypes.c_int64(R), ctypes.c_int64(C), qtype, nb, int(sym),
                              g, ctypes.c_void_p(0), ctypes.c_void_p(codes.data_ptr()), ctypes.c_void_p(scale.data_ptr()),
                              ctypes.c_void_p(zp.data_ptr()), ctypes.c_void_p(gs_out.data_ptr()))
    assert rc == 0
    geom = geom_of(name)
    mn, mx = O.minmax(w, geom)
    if qtype == O.FP4:
        gs = O.generate_gparam(float(w.float().min()), float(w.float().max()), dtype)
        assert gs.item() == gs_out.item()
        s_ref, _ = O.calculate_qparams(mn, mx, qtype, nb, sym, gs)
        want_scale = torch.tensor([O.lib().orc_f32_to_e4m3(float(v)) for v in s_ref.reshape(-1)], dtype=torch.uint8).reshape(R, G)
        assert_bits_equal(scale, want_scale, "e4m3 scale")
        q_ref = O.quantize(w, s_ref.to(dtype), torch.zeros(1), geom, qtype, nb, gs)
        assert_bits_equal(codes, q_ref, "fp4 nibbles")
        scale_T = s_ref.to(dtype)
    else:
        gs = None
        s_ref, z_ref = O.calculate_qparams(mn, mx, qtype, nb, sym)
        assert_bits_equal(scale, s_ref, "scale")
        if not sym:
            assert_bits_equal(zp, z_ref, "zp")
        q_ref = O.quantize(w, s_ref, z_ref if qtype == O.INT else torch.zeros(1), geom, qtype, nb)
        assert_bits_equal(codes, q_ref.view(torch.uint8), "codes")
        scale_T = s_ref
    # fake-quantize leg
    out = torch.empty_like(w)
    zpp = zp if (qtype == O.INT) else None
    gsp = gs if gs is not None else torch.ones(1)
    rc = hm.hm_fq_group(ctypes.c_void_p(w.data_ptr()), O._DT[dtype], ctypes.c_int64(R), ctypes.c_int64(C), qtype, nb, g,
                        ctypes.c_void_p(scale_T.data_ptr()), ctypes.c_void_p(zpp.data_ptr() if zpp is not None else 0), 1,
                        ctypes.c_void_p(gsp.data_ptr()), ctypes.c_void_p(out.data_ptr()))
    assert rc == 0
    want = O.fake_quantize(w, scale_T, zp if qtype == O.INT else torch.zeros(1), geom, qtype, nb, gs)
    assert_bits_equal(out, want, "fake_quantize")


def test_e4m3_and_gparam(hm):
    L = O.lib()
    for c in range(256):
        if (c & 0x7F) == 0x7F:
            continue
        assert hm.hm_e4m3_decode(c) == L.orc_e4m3_to_f32(c)
    g = torch.Generator().manual_seed(1)
    xs = torch.cat([torch.randn(5000, generator=g) * 50, torch.randn(5000, generator=g) * 0.02]).clamp(-448, 448)
    for v in xs.tolist():
        assert hm.hm_e4m3_encode(v) == L.orc_f32_to_e4m3(v)
    for dt, dtype in ((0, torch.bfloat16), (1, torch.float16), (2, torch.float32)):
        for a in torch.logspace(-10, 4, 500).to(dtype).float().tolist() + [0.0]:
            assert hm.hm_gparam(a, dt) == L.orc_gparam(-a, a, dt)
