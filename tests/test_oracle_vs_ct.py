"""CPU: cross-check the C oracle against LIVE compressed_tensors calls (skipped when the package is absent).
Complements the committed golden vectors with fresh seeds / shapes each format."""
import pytest
import torch

from oracle import ct_live as L
from oracle import oracle as O
from tests.util import FORMATS, assert_bits_equal, geom_of, synth_weight

pytestmark = pytest.mark.skipif(not L.available(), reason="compressed_tensors not importable")


@pytest.mark.parametrize("name", list(FORMATS))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_compress_live(name, dtype):
    fmt, args = L.format_args(name)
    _, qtype, nb, sym, *_ = FORMATS[name]
    for R, C, seed in ((16, 128, 1), (72, 640, 2), (200, 384, 3)):
        w = synth_weight(R, C, dtype, seed)
        ref = L.compress(w, fmt, args)
        got = O.compress(w, fmt, geom_of(name), nb, sym)
        assert set(ref) == set(got)
        for k in ref:
            assert_bits_equal(got[k], ref[k], f"{name}/{dtype}/{R}x{C}:{k}")


def test_nvfp4_supplied_global_scale_live():
    """fused q/k/v (gate/up) layers share min(global_scale): the compressor gets it from the state dict."""
    fmt, args = L.format_args("nvfp4")
    w = synth_weight(32, 256, torch.bfloat16, 7)
    gs = torch.tensor([L.global_scale(w).item() * 0.37])
    ref = L.compress(w, fmt, args, gs)
    got = O.compress(w, fmt, geom_of("nvfp4"), 4, True, gs)
    for k in ref:
        assert_bits_equal(got[k], ref[k], k)


def test_dequantize_roundtrip_live():
    from compressed_tensors.compressors.base import BaseCompressor
    from compressed_tensors.quantization import QuantizationScheme

    for name in ("int4_g128_asym", "int4_g32_sym", "fp8_block", "fp8_channel", "nvfp4"):
        fmt, args = L.format_args(name)
        _, qtype, nb, sym, *_ = FORMATS[name]
        w = synth_weight(128, 256, torch.bfloat16, 11)
        sd = L.compress(w, fmt, args)
        ref = BaseCompressor.get_value_from_registry(fmt).decompress(sd, QuantizationScheme(targets=["Linear"], weights=args))
        geom = geom_of(name)
        if fmt == "pack-quantized":
            q = O.unpack_from_int32(sd["weight_packed"], nb, w.shape)
            zp = None
            if not sym:
                zp = O.unpack_from_int32(sd["weight_zero_point"], nb, sd["weight_scale"].shape, 0)
            got = O.dequantize(q, sd["weight_scale"], zp, geom, qtype)
        elif fmt == "float-quantized":
            got = O.dequantize(sd["weight"], sd["weight_scale"], None, geom, qtype)
        else:
            vals = O.unpack_fp4_from_uint8(sd["weight_packed"], *w.shape)
            got = O.dequantize(vals, sd["weight_scale"].to(torch.bfloat16), None, geom, qtype, sd["weight_global_scale"],
                               out_dtype=torch.bfloat16)
        assert_bits_equal(got, ref["weight"], name)


@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g32_sym", "int4_channel_sym", "fp8_g128", "fp8_channel"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_mse_observer_live(name, dtype):
    """mse observer (O4): the restatement on the C oracle picks the same shrunk (min, max) per chunk as the restatement
    on the live compressed-tensors calls (calculate_qparams, TOKEN-patched fake_quantize)."""
    from oracle import llmc_restated as R

    _, args = L.format_args(name)
    _, qtype, nb, sym, *_ = FORMATS[name]
    w = synth_weight(24, 256, dtype, 5)
    rmn, rmx = L.mse_minmax(w, args)
    gmn, gmx = R.mse_minmax(w, geom_of(name), qtype, nb, sym)
    assert_bits_equal(gmn.reshape(rmn.shape), rmn, f"{name}/{dtype}: min")
    assert_bits_equal(gmx.reshape(rmx.shape), rmx, f"{name}/{dtype}: max")
    assert not torch.equal(rmn, torch.amin(L.flatten_weight(w, args), dim=(0, -1))), "the search must shrink some range"


@pytest.mark.parametrize("shape", [(4, 8), (7, 13), (16, 24), (130, 77)])
def test_pack_to_int32_out_of_range_codes_live(shape):
    """pack_to_int32 SUMS the shifted (code + offset) bytes (CT:compressors/pack_quantized/helpers.py:20-90): for int8 codes outside the
    4-bit range the carries spill into the neighbouring nibbles.  The oracle -- and, checked against it on the GPU, the CUDA kernels
    (tests/test_gpu_compress.py::test_flat_pack_unpack_fast_paths) -- must reproduce that, not an OR of masked nibbles."""
    from compressed_tensors.compressors.pack_quantized.helpers import pack_to_int32

    g = torch.Generator().manual_seed(shape[0] * 100 + shape[1])
    v = torch.randint(-128, 128, shape, generator=g, dtype=torch.int8)
    for dim in (1, 0):
        assert torch.equal(pack_to_int32(v, 4, packed_dim=dim), O.pack_to_int32(v, 4, dim)), (shape, dim)


@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g128_sym", "int4_channel_asym", "fp8_channel", "fp8_g32", "fp8_block", "nvfp4"])
def test_quantize_fake_quantize_foreign_qparams_live(name):
    """CT quantize / fake_quantize with qparams that are NOT the tensor's own statistics -- perturbed scales, a zero / 1e-35 / 3e5 scale,
    zero points incl. one outside the 4-bit range, an NVFP4 group scale that is not an e4m3 number: the oracle must agree with live
    compressed-tensors bit for bit, because the GPU tests of the supplied-qparams kernels (test_gpu_compress.py::*foreign*) are
    anchored on the oracle for exactly these cases."""
    from compressed_tensors.quantization.lifecycle.forward import fake_quantize, quantize

    fmt, args = L.format_args(name)
    _, qtype, nb, sym, strat, g, blk = FORMATS[name]
    geom = geom_of(name)
    rows, cols = (200, 392) if strat == O.BLOCK else (37, 768)
    gen = torch.Generator().manual_seed(3)
    w = synth_weight(rows, cols, torch.bfloat16, 91)
    mn, mx = O.minmax(w, geom)
    gs = O.generate_gparam(float(w.float().min()), float(w.float().max()), torch.bfloat16) if qtype == O.FP4 else None
    s, _ = O.calculate_qparams(mn, mx, qtype, nb, sym, gs)
    s = s.float() * (0.5 + 1.5 * torch.rand(s.shape, generator=gen))
    if qtype == O.FP4:
        s = s.clamp(max=448.0).to(torch.float8_e4m3fn).to(torch.bfloat16)
        s.view(-1)[1] = 0.3
    else:
        s = s.to(torch.bfloat16)
        if strat != O.BLOCK:  # block (0, 0) holds zero rows: 0 / 0 is a NaN whose payload is not comparable
            s.view(-1)[0] = 0.0
        s.view(-1)[1] = 1e-35
        s.view(-1)[-1] = 3.0e5
    if qtype == O.INT:
        zp = torch.randint(-8, 8, s.shape, generator=gen, dtype=torch.int8) if not sym else torch.zeros(s.shape, dtype=torch.int8)
        if not sym:
            zp.view(-1)[2] = 40
    else:
        zp = torch.zeros(s.shape, dtype=torch.float8_e4m3fn)
    q_live = quantize(w, s, zp, args, global_scale=gs)
    fq_live = fake_quantize(w, s, zp, args, global_scale=gs)
    q_o = O.quantize(w, s, zp, geom, qtype, nb, gs)
    if qtype == O.INT:
        assert torch.equal(q_live.to(torch.int8), q_o), name
    elif qtype == O.FP8:
        assert_bits_equal(q_live.to(torch.float8_e4m3fn), q_o, name)
    else:
        vals = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])[(q_o & 7).long()] * torch.where((q_o & 8) > 0, -1.0, 1.0)
        assert_bits_equal(q_live, vals.to(torch.bfloat16), name)
    assert_bits_equal(fq_live, O.fake_quantize(w, s, zp, geom, qtype, nb, gs), name)


@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g128_sym", "int4_channel_asym"])
def test_dequantize_int8_codes_foreign_qparams_live(name):
    """CT dequantize on un-packed int8 codes over the FULL int8 range with foreign scales / zero points (and without a zero point):
    anchors the GPU test of dequant_int8_fast_kernel."""
    from compressed_tensors.quantization.lifecycle.forward import dequantize

    fmt, args = L.format_args(name)
    _, qtype, nb, sym, strat, gsz, blk = FORMATS[name]
    geom = geom_of(name)
    g = torch.Generator().manual_seed(9)
    rows, cols = 37, 768
    w = synth_weight(rows, cols, torch.bfloat16, 5)
    s, _ = O.calculate_qparams(*O.minmax(w, geom), qtype, nb, sym)
    s = (s.float() * (0.5 + 1.5 * torch.rand(s.shape, generator=g))).to(torch.bfloat16)
    zp = torch.randint(-8, 8, s.shape, generator=g, dtype=torch.int8)
    codes = torch.randint(-128, 128, (rows, cols), generator=g, dtype=torch.int8)
    for z in (zp, None):
        assert_bits_equal(dequantize(codes, s, z, args=args), O.dequantize(codes, s, z, geom, qtype), f"{name} zp={z is not None}")


def test_pack_fp4_off_grid_values_live():
    """pack_fp4_to_uint8 on values that are NOT e2m1 grid points (first minimum of |x| - grid evaluated in bf16, sign bit from the
    input, magnitudes beyond 6): anchors the slow path of pack_fp4_flat_kernel."""
    from compressed_tensors.compressors.nvfp4.helpers import pack_fp4_to_uint8

    g = torch.Generator().manual_seed(9)
    off = (torch.randn(16, 64, generator=g) * 3).to(torch.bfloat16)
    off[0, :6] = torch.tensor([-0.0, 7.5, -100.0, 0.25, 0.75, -0.25], dtype=torch.bfloat16)
    assert torch.equal(pack_fp4_to_uint8(off), O.pack_fp4_to_uint8(off))
