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
