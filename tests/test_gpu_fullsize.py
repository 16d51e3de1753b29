"""GPU: BASELINE.json-size matrices.  The oracle is too slow to re-run at these sizes, so parity rests on
size-independent properties: (1) row-sliced agreement with the oracle on a sample of rows (group formats are
row independent), (2) decompress(compress(w)) == fake_quantize(w) bit-for-bit, (3) idempotence of the codes,
(4) a checksum of checksums that is independent of how the launch is tiled (stack vs. loop)."""
import pytest
import torch

from oracle import oracle as O
from tests.test_gpu_compress import Args
from tests.util import FORMATS, assert_bits_equal, geom_of, synth_weight

pytestmark = pytest.mark.gpu


def _w(R, C, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(R, C, generator=g, device="cuda") * 0.02
    w[:, ::997] *= 20
    return w.to(torch.bfloat16)


@pytest.mark.parametrize("name,shape", [("int4_g128_asym", (2560, 9728)), ("int4_g32_sym", (9728, 2560)),
                                        ("fp8_block", (4096, 2560)), ("fp8_block", (2560 + 64, 9728)),
                                        ("fp8_channel", (2560, 4096)), ("nvfp4", (768, 2048))])
def test_fullsize_properties(name, shape):
    from quantizers_b200 import ops

    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    args = Args(name)
    R, C = shape
    w = _w(R, C, 1234)
    sd = ops.compress_weight(w, args)
    # (1) sampled rows vs the oracle (row-independent strategies only)
    if strat in (O.GROUP, O.CHANNEL) and qtype != O.FP4:
        rows = torch.tensor([0, 1, R // 3, R // 2, R - 2, R - 1])
        rows8 = (rows // 8 * 8).unique()
        idx = (rows8[:, None] + torch.arange(8)[None, :]).reshape(-1)
        sub = w[idx.cuda()].cpu()
        want = O.compress(sub, fmt, geom_of(name), nb, sym)
        key = "weight_packed" if qtype == O.INT else "weight"
        assert_bits_equal(sd[key][idx.cuda()], want[key], f"{name}: sampled rows")
        assert_bits_equal(sd["weight_scale"][idx.cuda()], want["weight_scale"], f"{name}: sampled scales")
    # (2) round trip == fake quantize
    if qtype == O.INT:
        q = ops.unpack_from_int32(sd["weight_packed"], nb, (R, C))
        zp = None
        if not sym:
            zp = ops.unpack_from_int32(sd["weight_zero_point"], nb, sd["weight_scale"].shape, 0)
        deq = ops.dequantize(q, sd["weight_scale"], zp, args)
        fq = ops.fake_quantize(w, sd["weight_scale"], zp if zp is not None else torch.zeros_like(sd["weight_scale"], dtype=torch.int8), args)
    elif qtype == O.FP8:
        deq = ops.dequantize(sd["weight"], sd["weight_scale"], None, args)
        fq = ops.fake_quantize(w, sd["weight_scale"], torch.zeros(1, device="cuda"), args)
    else:
        vals = ops.unpack_fp4_from_uint8(sd["weight_packed"], R, C)
        sT = sd["weight_scale"].to(torch.bfloat16)
        deq = ops.dequantize(vals, sT, None, args, dtype=torch.bfloat16, global_scale=sd["weight_global_scale"])
        fq = ops.fake_quantize(w, sT, torch.zeros(1, device="cuda"), args, global_scale=sd["weight_global_scale"])
    # -0.0 vs +0.0 is the only admissible difference between the two routes (storage codes drop the sign of INT zeros)
    assert torch.equal(deq.float(), fq.float())
    # (3) idempotence: quantizing the de-quantized weight with the same qparams reproduces the codes
    if qtype == O.INT:
        q2 = ops.quantize_pack(deq, sd["weight_scale"], zp, args)
        assert torch.equal(q2, sd["weight_packed"])
    # (FP8/FP4 are not idempotent through T: bf16(q*s)/s loses bits of the 4-bit-significand x 8-bit-significand product)
    # (4) tiling independence: the same matrix inside a 3-stack gives the same bytes
    st = ops.compress_weight(torch.stack([w, w.flip(0), w]), args)
    key = "weight_packed" if "weight_packed" in sd else "weight"
    assert torch.equal(st[key][0].view(torch.uint8), sd[key].view(torch.uint8))
    assert torch.equal(st[key][2].view(torch.uint8), sd[key].view(torch.uint8))
    assert torch.equal(st["weight_scale"][2].view(torch.uint8), sd["weight_scale"].view(torch.uint8))
