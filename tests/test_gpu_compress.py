"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle, the committed golden vectors
and, when importable, live compressed_tensors.  Bit-exact everywhere (integer / byte / index work and scales)."""
import numpy as np
import pytest
import torch

from oracle import ct_live as L
from oracle import oracle as O
from tests.util import FORMATS, assert_bits_equal, from_bits, geom_of, golden_files, load_golden, synth_weight

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float16, torch.float32]


class Args:
    """Minimal QuantizationArgs look-alike (the ops accept any object with these attributes)."""

    def __init__(self, name):
        fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
        self.num_bits = nb
        self.type = "int" if qtype == O.INT else "float"
        self.symmetric = sym
        self.strategy = {O.TENSOR: "tensor", O.CHANNEL: "channel", O.GROUP: "tensor_group" if qtype == O.FP4 else "group",
                         O.BLOCK: "block"}[strat]
        self.group_size = g or None
        self.block_structure = list(blk) if blk else None
        self.zp_dtype = torch.int8 if qtype == O.INT else torch.float8_e4m3fn


def _cmp_sd(got, want, what):
    assert set(got) == set(want), f"{what}: keys {sorted(got)} != {sorted(want)}"
    for k in want:
        w = want[k]
        if isinstance(w, np.ndarray):
            assert_bits_equal(got[k].reshape(w.shape), w, f"{what}:{k}")
        else:
            assert_bits_equal(got[k].reshape(w.shape), w, f"{what}:{k}")


@pytest.mark.parametrize("fname", golden_files())
def test_fused_compress_matches_golden(fname):
    from quantizers_b200 import ops

    name, dtype, z = load_golden(fname)
    w = from_bits(z["w"], dtype).cuda()
    got = ops.compress_weight(w, Args(name))
    want = {k[3:]: v for k, v in z.items() if k.startswith("sd_")}
    _cmp_sd(got, want, name)


@pytest.mark.parametrize("name", list(FORMATS))
@pytest.mark.parametrize("dtype", DTYPES)
def test_fused_compress_matches_oracle(name, dtype):
    from quantizers_b200 import ops

    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    for R, C, seed in ((8, 128, 1), (77, 1280, 2), (200, 384, 3), (256, 2304, 4)):
        w = synth_weight(R, C, dtype, seed)
        want = O.compress(w, fmt, geom_of(name), nb, sym)
        got = ops.compress_weight(w.cuda(), Args(name))
        _cmp_sd(got, want, f"{name}/{dtype}/{R}x{C}")


@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g32_sym", "fp8_block", "fp8_g32", "nvfp4", "int4_channel_asym"])
def test_fused_compress_expert_stack(name):
    """[E, rows, cols] stacks (MoE experts) == per-expert results; rows not a multiple of 8 exercises zp row packing."""
    from quantizers_b200 import ops

    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    E, R, C = 5, 36, 512
    ws = [synth_weight(R, C, torch.bfloat16, 100 + e) for e in range(E)]
    got = ops.compress_weight(torch.stack(ws).cuda(), Args(name))
    for e in range(E):
        want = O.compress(ws[e], fmt, geom_of(name), nb, sym)
        for k in want:
            if k == "weight_shape":
                continue
            assert_bits_equal(got[k][e].reshape(want[k].shape), want[k], f"{name}[{e}]:{k}")


def test_nvfp4_supplied_global_scale():
    from quantizers_b200 import ops

    w = synth_weight(64, 512, torch.bfloat16, 9)
    gs = torch.tensor([37.25])
    want = O.compress(w, "nvfp4-pack-quantized", geom_of("nvfp4"), 4, True, gs)
    got = ops.compress_weight(w.cuda(), Args("nvfp4"), global_scale=gs.cuda())
    _cmp_sd(got, want, "nvfp4 gs")


@pytest.mark.parametrize("persistent", ["0", "1", "fallback", "nt1", "nt8"])
@pytest.mark.parametrize("E,R,C,span", [(6, 36, 512, 2), (64, 768, 2048, 2), (9, 200, 1040, 3), (3, 2048, 768, 1), (320, 256, 2048, 2)])
def test_nvfp4_fused_sibling_global_scale(E, R, C, span, persistent, monkeypatch):
    """gate/up siblings stacked next to each other share min(global_scale) (LLMC update_fused_layer_weight_global_scales); both
    schedulings of the single-launch kernel (one CTA per item with the |max| pass running ahead; persistent warp-specialised CTA
    per SM, B200Q_FP4_PERSISTENT=1) must equal the oracle run with that shared scale.  "fallback": every compress CTA takes the
    never-observed branch that reduces the span itself (the launch must not depend on CTA dispatch order); "nt1" / "nt8": other
    tiles-per-CTA counts than the default (whole-tile fast path vs the ragged last tile, the CTA-index decode)."""
    from quantizers_b200 import ops

    monkeypatch.setenv("B200Q_FP4_PERSISTENT", "1" if persistent == "1" else "0")
    if persistent == "fallback":
        monkeypatch.setenv("B200Q_FP4_FORCE_FALLBACK", "1")
    if persistent.startswith("nt"):
        monkeypatch.setenv("B200Q_FP4_NT", persistent[2:])

    ws = [(synth_weight(R, C, torch.bfloat16, 300 + e, edge=R >= 64) * (1.0 + 0.37 * e)).to(torch.bfloat16) for e in range(E)]
    got = ops.compress_weight(torch.stack(ws).cuda(), Args("nvfp4"), fuse_span=span)
    for s0 in range(0, E, span):
        gs = min(float(O.generate_gparam(float(ws[e].float().min()), float(ws[e].float().max()), torch.bfloat16)) for e in range(s0, s0 + span))
        for e in range(s0, s0 + span):
            want = O.compress(ws[e], "nvfp4-pack-quantized", geom_of("nvfp4"), 4, True, torch.tensor([gs]))
            for k in want:
                assert_bits_equal(got[k][e].reshape(want[k].shape), want[k], f"nvfp4 fused[{e}]:{k}")


@pytest.mark.parametrize("name", list(FORMATS))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_unfused_ops_match_oracle(name, dtype):
    """observe_minmax -> calculate_qparams -> quantize / quantize_pack / fake_quantize / dequantize, each vs the oracle."""
    from quantizers_b200 import ops

    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    geom = geom_of(name)
    args = Args(name)
    R, C = (200, 384) if strat == O.BLOCK else (40, 768)
    w = synth_weight(R, C, dtype, 17)
    wd = w.cuda()
    gs = O.generate_gparam(float(w.float().min()), float(w.float().max()), dtype) if qtype == O.FP4 else None
    gsd = gs.cuda() if gs is not None else None
    # observer
    if strat != O.TENSOR or True:
        mn_o, mx_o = O.minmax(w, geom)
        mn, mx = ops.observe_minmax(wd, args)
        if strat != O.BLOCK or (R % 128 == 0 and C % 128 == 0):
            assert_bits_equal(mn.reshape(mn_o.shape), mn_o, "min")
            assert_bits_equal(mx.reshape(mx_o.shape), mx_o, "max")
    if qtype == O.FP4:
        assert ops.observe_global_scale(wd).item() == gs.item()
    # qparams
    s_o, z_o = O.calculate_qparams(mn_o, mx_o, qtype, nb, sym, gs)
    s, zp = ops.calculate_qparams(mn_o.cuda(), mx_o.cuda(), args, gsd)
    assert_bits_equal(s.reshape(s_o.shape), s_o, "scale")
    if qtype == O.INT:
        assert_bits_equal(zp.reshape(z_o.shape), z_o, "zp")
    sT = s_o.to(dtype)
    zarg = z_o if qtype == O.INT else torch.zeros(s_o.shape, dtype=torch.float8_e4m3fn)
    # quantize
    q_o = O.quantize(w, sT, zarg, geom, qtype, nb, gs)
    if qtype == O.FP4:
        q = ops.quantize(wd, sT.cuda(), zarg.cuda(), args, global_scale=gsd)
        vals = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])[(q_o & 7).long()] * torch.where((q_o & 8) > 0, -1.0, 1.0)
        assert_bits_equal(q, vals.to(dtype), "fp4 values")
        packed = ops.pack_fp4_to_uint8(q)
        assert_bits_equal(packed, (q_o[:, 0::2] | (q_o[:, 1::2] << 4)), "pack_fp4")
        assert_bits_equal(ops.unpack_fp4_from_uint8(packed, R, C, dtype), vals.to(dtype), "unpack_fp4")
    else:
        q = ops.quantize(wd, sT.cuda(), zarg.cuda(), args, dtype=torch.int8 if qtype == O.INT else torch.float8_e4m3fn)
        assert_bits_equal(q, q_o, "codes")
    # quantize_pack == compress given qparams
    qp = ops.quantize_pack(wd, sT.cuda(), zarg.cuda(), args, global_scale=gsd)
    if qtype == O.INT:
        assert_bits_equal(qp, O.pack_to_int32(q_o, nb), "quantize_pack")
        assert_bits_equal(ops.unpack_from_int32(qp, nb, (R, C)), q_o, "unpack")
    elif qtype == O.FP8:
        assert_bits_equal(qp, q_o, "quantize_pack fp8")
    else:
        assert_bits_equal(qp, (q_o[:, 0::2] | (q_o[:, 1::2] << 4)), "quantize_pack fp4")
    # fake_quantize
    fq_o = O.fake_quantize(w, sT, zarg, geom, qtype, nb, gs)
    fq = ops.fake_quantize(wd, sT.cuda(), zarg.cuda(), args, global_scale=gsd)
    assert_bits_equal(fq, fq_o, "fake_quantize")
    # dequantize
    if qtype == O.FP4:
        dq_o = O.dequantize(vals.to(dtype), sT, None, geom, qtype, gs, out_dtype=dtype)
        dq = ops.dequantize(vals.to(dtype).cuda(), sT.cuda(), None, args, dtype=dtype, global_scale=gsd)
    else:
        dq_o = O.dequantize(q_o, sT, z_o if qtype == O.INT else None, geom, qtype)
        dq = ops.dequantize(q_o.cuda(), sT.cuda(), z_o.cuda() if qtype == O.INT else None, args)
    assert_bits_equal(dq, dq_o, "dequantize")


def test_pack_zero_point_rows_and_ragged():
    from quantizers_b200 import ops

    g = torch.Generator().manual_seed(3)
    for R, C in [(1, 1), (7, 13), (9, 8), (16, 24), (130, 77)]:
        v = torch.randint(-8, 8, (R, C), generator=g, dtype=torch.int8)
        for dim in (0, 1):
            p = ops.pack_to_int32(v.cuda(), 4, dim)
            assert_bits_equal(p, O.pack_to_int32(v, 4, dim), f"pack {R}x{C} dim{dim}")
            assert torch.equal(ops.unpack_from_int32(p, 4, v.shape, dim).cpu(), v)
    with pytest.raises(ValueError):
        ops.pack_to_int32(torch.zeros(2, 2, dtype=torch.int32).cuda(), 4)
    with pytest.raises(ValueError):
        ops.pack_fp4_to_uint8(torch.zeros(2, 3, dtype=torch.bfloat16).cuda())


def test_error_behaviour_matches_reference():
    from quantizers_b200 import ops

    a = Args("int4_g128_sym")
    with pytest.raises(ValueError, match="divisble"):
        ops.compress_weight(torch.zeros(8, 200, dtype=torch.bfloat16).cuda(), a)
    w = torch.zeros(8, 200, dtype=torch.bfloat16).cuda()
    with pytest.raises(ValueError, match="divisble"):
        ops.fake_quantize(w, torch.ones(8, 2, dtype=torch.bfloat16).cuda(), None, a)
    # empty input: nothing to do, nothing launched
    e = ops.fake_quantize(torch.zeros(0, 128, dtype=torch.bfloat16).cuda(), torch.ones(0, 1, dtype=torch.bfloat16).cuda(), None, a)
    assert e.numel() == 0


@pytest.mark.skipif(not L.available(), reason="compressed_tensors not importable")
@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g32_sym", "int4_channel_asym", "fp8_block", "fp8_channel", "fp8_g32", "nvfp4"])
def test_registered_compressors_match_live_ct(name):
    """The drop-in seam: BaseCompressor registry override vs the stock compressor on the same state dict, and
    decompress round trip through the CUDA unpack/dequantize kernels vs CT's."""
    from compressed_tensors.compressors.base import BaseCompressor
    from compressed_tensors.quantization import QuantizationScheme

    import quantizers_b200.patch as P

    fmt, args = L.format_args(name)
    scheme = QuantizationScheme(targets=["Linear"], weights=args)
    w = synth_weight(136, 640, torch.bfloat16, 23)
    gs = L.global_scale(w) if args.strategy == "tensor_group" else None
    scale, zp = L.weight_qparams(w, args, gs)
    sd = {"weight": w, "weight_scale": scale, "weight_zero_point": zp}
    if gs is not None:
        sd["weight_global_scale"] = gs
    ref = BaseCompressor.get_value_from_registry(fmt).compress(sd, scheme)
    ref_dec = BaseCompressor.get_value_from_registry(fmt).decompress(ref, scheme)
    sd_cuda = {k: v.cuda() for k, v in sd.items()}
    keys_before = set(sd_cuda)
    with P.patch():
        comp = BaseCompressor.get_value_from_registry(fmt)
        assert comp.__name__.startswith("B200")
        got = comp.compress(sd_cuda, scheme)
        got_dec = comp.decompress(got, scheme)
    assert set(sd_cuda) == keys_before  # input dict not modified
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].dtype == ref[k].dtype, k
        assert_bits_equal(got[k], ref[k], f"{name}:{k}")
    assert_bits_equal(got_dec["weight"], ref_dec["weight"], f"{name}:decompress")
    assert BaseCompressor.get_value_from_registry(fmt).__name__.startswith("B200") is False


@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g128_sym", "int4_g32_sym", "int4_g32_asym", "fp8_block", "fp8_g32",
                                  "fp8_channel", "nvfp4"])
def test_generic_kernels_still_match_oracle_bf16(name, monkeypatch):
    """bf16 normally takes the issue-tuned fast kernels; B200Q_DISABLE_FAST=1 forces the generic templates."""
    from quantizers_b200 import ops

    monkeypatch.setenv("B200Q_DISABLE_FAST", "1")
    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    w = synth_weight(72, 1280, torch.bfloat16, 31)
    want = O.compress(w, fmt, geom_of(name), nb, sym)
    got = ops.compress_weight(w.cuda(), Args(name))
    _cmp_sd(got, want, name)


def test_fast_int4_boundary_cases():
    """Adversarial inputs for the reciprocal-multiply fast path: quotients that land on / next to bf16 rounding
    boundaries, huge / tiny scales (exact-path fallback), exact zeros, and a dense sweep of all bf16 magnitudes."""
    from quantizers_b200 import ops

    rows = []
    # every finite positive bf16 bit pattern (and negatives) as data, in groups whose absmax is set by column 0
    allv = torch.arange(0x0001, 0x7F80, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float()
    for scale_max in (1.0, 0.0371, 7.5, 3.0e-3, 1.0e30, 1.0e-30, 448.0):
        v = allv[(allv <= scale_max)][-(127 * 64):]
        v = v[: (v.numel() // 127) * 127]
        blk = v.reshape(-1, 127)
        blk = torch.cat([torch.full((blk.shape[0], 1), scale_max), blk * torch.where(torch.arange(127) % 2 == 0, 1.0, -1.0)], dim=1)
        rows.append(blk)
    w = torch.cat(rows).to(torch.bfloat16)
    w = w[: (w.shape[0] // 8) * 8]
    for name in ("int4_g128_asym", "int4_g128_sym", "int4_g32_sym", "int4_g32_asym", "fp8_g32", "fp8_g128", "fp8_block", "nvfp4"):
        fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
        want = O.compress(w, fmt, geom_of(name), nb, sym)
        got = ops.compress_weight(w.cuda(), Args(name))
        _cmp_sd(got, want, name)
    # NVFP4 thresholds: quotients exactly on / one ulp around the e2m1 rounding points, scale fixed by column 0 = 6.0
    th = torch.tensor([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, 0.0])
    th_bits = th.to(torch.bfloat16).view(torch.int16).int()
    cols = [torch.full((15,), 6.0)]
    for d in (-1, 0, 1):
        v = (th_bits + d).clamp(min=0).to(torch.int16).view(torch.bfloat16).float()
        cols += [v, -v]
    blk16 = torch.stack(cols + [torch.zeros(15)] * (16 - len(cols)), dim=1)  # [15, 16]
    w2 = blk16.repeat(8, 8).to(torch.bfloat16)  # [120, 128]
    want = O.compress(w2, "nvfp4-pack-quantized", geom_of("nvfp4"), 4, True)
    got = ops.compress_weight(w2.cuda(), Args("nvfp4"))
    _cmp_sd(got, want, "nvfp4 thresholds")


@pytest.mark.parametrize("name", ["fp8_channel", "int4_channel_sym", "int4_channel_asym", "int8_channel_sym"])
def test_fast_channel_kernel_team_sizes_and_boundaries(name):
    """The bf16 CHANNEL kernel gives a row to a team of 1 / 2 / 4 / 8 warps depending on its length (rows longer than 16384
    fall back to the generic kernel): every team size, ragged row counts (last CTA partly empty), stacked matrices (zero-point
    words packed down the rows of EACH matrix), plus adversarial rows -- all bf16 magnitudes against scales from 1e-30 to 1e30,
    all-zero / all-negative / all-positive rows, signed zeros."""
    from quantizers_b200 import ops

    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    for R, C, seed in ((13, 2048, 1), (11, 2056, 2), (9, 4096, 3), (7, 4104, 4), (5, 8192, 5), (3, 8200, 6), (3, 16384, 7), (2, 16392, 8)):
        w = synth_weight(R, C, torch.bfloat16, seed)
        _cmp_sd(ops.compress_weight(w.cuda(), Args(name)), O.compress(w, fmt, geom_of(name), nb, sym), f"{name}/{R}x{C}")
    # stacked launch: [3, 21, 2560] -> per-matrix results, zero-point rows do not bleed across matrices
    ws = [synth_weight(21, 2560, torch.bfloat16, 40 + i) for i in range(3)]
    got = ops.compress_weight(torch.stack(ws).cuda(), Args(name))
    for i, w in enumerate(ws):
        want = O.compress(w, fmt, geom_of(name), nb, sym)
        for k, v in want.items():
            if k != "weight_shape":
                assert_bits_equal(got[k][i].reshape(v.shape), v, f"{name}: stacked[{i}].{k}")
    allv = torch.arange(0x0001, 0x7F80, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float()
    rows = []
    for scale_max in (1.0, 0.0371, 7.5, 448.0, 3.0e-3, 1.0e30, 1.0e-30):
        v = allv[(allv <= scale_max)][-(255 * 24):]
        v = v[: (v.numel() // 255) * 255].reshape(-1, 255)
        rows.append(torch.cat([torch.full((v.shape[0], 1), scale_max), v * torch.where(torch.arange(255) % 2 == 0, 1.0, -1.0)], dim=1))
    special = torch.zeros(6, 256)
    special[1] = -torch.rand(256) - 0.1            # all negative
    special[2] = torch.rand(256) + 0.1             # all positive
    special[3, ::2] = -0.0                         # signed zeros only
    special[4, 5] = 3.0e38                         # one huge element
    special[5, 7] = -1.0e-38                       # one subnormal-scale element
    w = torch.cat(rows + [special]).to(torch.bfloat16)
    _cmp_sd(ops.compress_weight(w.cuda(), Args(name)), O.compress(w, fmt, geom_of(name), nb, sym), f"{name}/adversarial")


@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g32_sym", "int4_channel_asym", "int4_channel_sym", "int8_g128_sym", "nvfp4"])
def test_fused_decompress_matches_two_step_oracle(name):
    """§8f rank 1: packed codes + qparams -> weights in one pass == the oracle's unpack followed by dequantize (which the CPU
    suite pins against live compressed-tensors' Compressor.decompress), incl. a stacked [E, rows, cols] launch."""
    from quantizers_b200 import ops

    fmt, qtype, nb, sym, *_ = FORMATS[name]
    geom = geom_of(name)
    ws = [synth_weight(72, 256, torch.bfloat16, 60 + e) for e in range(3)]
    sds = [O.compress(w, fmt, geom, nb, sym) for w in ws]
    wants = []
    for w, sd in zip(ws, sds):
        if fmt == "pack-quantized":
            q = O.unpack_from_int32(sd["weight_packed"], nb, w.shape)
            zp = None if sym else O.unpack_from_int32(sd["weight_zero_point"], nb, sd["weight_scale"].shape, 0)
            wants.append(O.dequantize(q, sd["weight_scale"], zp, geom, qtype))
        else:
            vals = O.unpack_fp4_from_uint8(sd["weight_packed"], *w.shape)
            wants.append(O.dequantize(vals, sd["weight_scale"].to(torch.bfloat16), None, geom, qtype, sd["weight_global_scale"], out_dtype=torch.bfloat16))
    stack = lambda k: torch.stack([sd[k] for sd in sds]).cuda()
    if fmt == "pack-quantized":
        got = ops.decompress_int_packed(stack("weight_packed"), stack("weight_scale"), None if sym else stack("weight_zero_point"), (72, 256), Args(name))
    else:
        got = ops.decompress_nvfp4(stack("weight_packed"), stack("weight_scale"), stack("weight_global_scale"))
    for e, want in enumerate(wants):
        assert_bits_equal(got[e], want, f"{name}[{e}]")


@pytest.mark.parametrize("name", list(FORMATS))
def test_fused_compress_random_ragged_shapes(name):
    """Seeded sweep over awkward shapes: single rows, a single group / block per row, row counts that leave the last CTA, warp
    tile or zero-point word partly empty, column counts just above a tile boundary, stacks of 1-5 matrices, value scales from
    1e-6 to 1e4 -- every output tensor bit-for-bit against the oracle."""
    import random
    import zlib

    from quantizers_b200 import ops

    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    unit = {O.GROUP: g, O.BLOCK: 8, O.CHANNEL: 8, O.TENSOR: 8}[strat]
    if qtype == O.FP4:
        unit = 16
    rng = random.Random(zlib.crc32(name.encode()))  # str hashes are salted per process
    cases = [(1, unit), (1, unit * 3), (2, unit * 33), (7, unit), (9, max(unit, 136) // unit * unit)]
    while len(cases) < 22:
        cases.append((rng.choice([1, 3, 8, 15, 17, 64, 129, 200]), unit * rng.choice([1, 2, 5, 8, 17, 32, 33, 40])))
    for i, (R, C) in enumerate(cases):
        if R * C > 600_000:
            continue
        n = rng.choice([1, 1, 2, 5])
        scale = 10.0 ** rng.uniform(-6, 4)
        ws = [(synth_weight(R, C, torch.float32, 500 + 7 * i + e, edge=(i % 3 == 0)) * scale).to(torch.bfloat16) for e in range(n)]
        got = ops.compress_weight(torch.stack(ws).cuda() if n > 1 else ws[0].cuda(), Args(name))
        for e, w in enumerate(ws):
            want = O.compress(w, fmt, geom_of(name), nb, sym)
            for k, v in want.items():
                if k == "weight_shape":
                    continue
                gk = got[k][e] if n > 1 else got[k]
                assert_bits_equal(gk.reshape(v.shape), v, f"{name} case {i}: {n} x [{R}, {C}] scale {scale:.2e} [{e}].{k}")


@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g128_sym", "int4_g32_sym", "int4_g32_asym", "fp8_g32", "fp8_g128"])
@pytest.mark.parametrize("rows,cols", [(37, 768), (64, 2560), (3, 128)])
def test_quantize_pack_with_foreign_qparams(name, rows, cols):
    """Compressor.compress's arithmetic with qparams that did NOT come from this weight's min/max (an EMA / MSE observer, a
    checkpoint): perturbed scales, random zero points, a zero scale and a huge one.  The bf16 fast path (TMA kernel, SUPPLIED
    mode) must use exactly what it is given; partial last tiles and stacked matrices included."""
    from quantizers_b200 import ops

    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    geom, args = geom_of(name), Args(name)
    gen = torch.Generator().manual_seed(rows * 131 + cols)
    ws = [synth_weight(rows, cols, torch.bfloat16, 50 + i) for i in range(2)]
    G = cols // g
    packs, scales, zps = [], [], []
    for w in ws:
        mn, mx = O.minmax(w, geom)
        s, z = O.calculate_qparams(mn, mx, qtype, nb, sym)
        s = (s.float() * (0.5 + 1.5 * torch.rand(s.shape, generator=gen))).to(torch.bfloat16)
        s.view(-1)[0] = 0.0          # x / 0
        s.view(-1)[-1] = 3.0e5       # outside the reciprocal bracket's safe range
        if G > 2:
            s.view(-1)[1] = 1e-35
        z = torch.randint(-8, 8, s.shape, generator=gen, dtype=torch.int8) if (qtype == O.INT and not sym) else None
        zarg = z if z is not None else (None if qtype == O.INT else torch.zeros(s.shape, dtype=torch.float8_e4m3fn))
        q_o = O.quantize(w, s, zarg, geom, qtype, nb)
        packs.append(O.pack_to_int32(q_o, nb) if qtype == O.INT else q_o)
        scales.append(s)
        zps.append(zarg)
    zp_stack = None if zps[0] is None else torch.stack(zps).cuda()
    got = ops.quantize_pack(torch.stack(ws).cuda(), torch.stack(scales).cuda(), zp_stack, args)
    for i in range(2):
        assert_bits_equal(got[i], packs[i], f"{name}[{i}]")


@pytest.mark.parametrize("name", ["int4_g128_asym", "int4_g128_sym", "int4_g32_asym", "int4_channel_asym", "int4_channel_sym",
                                  "fp8_channel", "fp8_g32", "fp8_block", "fp8_tensor"])
@pytest.mark.parametrize("rows,cols", [(37, 768), (200, 384 + 8), (5, 128)])
def test_quantize_fake_quantize_with_foreign_qparams(name, rows, cols):
    """CT quantize / fake_quantize (bf16 fast path, elementwise_fast.cu) with qparams that are not this tensor's own statistics:
    perturbed scales, random zero points (one outside [-8, 7]), zero / tiny / huge scales, the -0.0 row of synth_weight (a negative
    value rounding to zero must dequantize to -0.0), ragged blocks and rows that are not a multiple of the 1024-column batch."""
    from quantizers_b200 import ops

    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    if strat == O.GROUP and cols % g:
        pytest.skip("columns not divisible by the group")
    geom, args = geom_of(name), Args(name)
    gen = torch.Generator().manual_seed(rows * 17 + cols)
    w = synth_weight(rows, cols, torch.bfloat16, 91)
    mn, mx = O.minmax(w, geom)
    s, z = O.calculate_qparams(mn, mx, qtype, nb, sym)
    s = (s.float() * (0.5 + 1.5 * torch.rand(s.shape, generator=gen))).to(torch.bfloat16)
    if s.numel() > 3:
        if strat != O.BLOCK:  # block (0, 0) holds synth_weight's zero rows: 0 / 0 is a NaN whose SIGN differs between x86 and the GPU
            s.view(-1)[0] = 0.0
        s.view(-1)[1] = 1e-35
        s.view(-1)[-1] = 3.0e5
    for use_zp in ((True, False) if qtype == O.INT else (True,)):
        if qtype == O.INT:
            zarg = None
            if use_zp:
                zarg = torch.randint(-8, 8, s.shape, generator=gen, dtype=torch.int8) if not sym else torch.zeros(s.shape, dtype=torch.int8)
                if not sym and zarg.numel() > 2:
                    zarg.view(-1)[2] = 40
        else:
            zarg = torch.zeros(s.shape, dtype=torch.float8_e4m3fn)
        q_o = O.quantize(w, s, zarg, geom, qtype, nb)
        q = ops.quantize(w.cuda(), s.cuda(), None if zarg is None else zarg.cuda(), args,
                         dtype=torch.int8 if qtype == O.INT else torch.float8_e4m3fn)
        assert_bits_equal(q, q_o, f"{name} quantize zp={use_zp}")
        fq_o = O.fake_quantize(w, s, zarg, geom, qtype, nb)
        fq = ops.fake_quantize(w.cuda(), s.cuda(), None if zarg is None else zarg.cuda(), args)
        assert_bits_equal(fq, fq_o, f"{name} fake_quantize zp={use_zp}")
        if qtype == O.INT:  # un-packed int8 codes -> bf16 (full int8 range, not only the 4-bit codes)
            codes = torch.randint(-128, 128, (rows, cols), generator=gen, dtype=torch.int8)
            dq_o = O.dequantize(codes, s, zarg, geom, qtype)
            dq = ops.dequantize(codes.cuda(), s.cuda(), None if zarg is None else zarg.cuda(), args)
            assert_bits_equal(dq, dq_o, f"{name} dequantize zp={use_zp}")


@pytest.mark.parametrize("rows,cols", [(37, 768), (64, 2560), (3, 16)])
def test_nvfp4_quantize_pack_with_foreign_scales(rows, cols):
    """nvfp4 Compressor.compress's arithmetic with the module's own weight_scale / weight_global_scale: group scales that are NOT what
    this weight's |max| would give (neighbouring e4m3 codes), a zero scale, a value that is not an e4m3 number, stacked matrices."""
    from quantizers_b200 import ops

    geom, args = geom_of("nvfp4"), Args("nvfp4")
    gen = torch.Generator().manual_seed(rows + cols)
    ws = [synth_weight(rows, cols, torch.bfloat16, 70 + i) for i in range(2)]
    gs = O.generate_gparam(min(float(w.float().min()) for w in ws), max(float(w.float().max()) for w in ws), torch.bfloat16)
    packs, scales = [], []
    for w in ws:
        mn, mx = O.minmax(w, geom)
        s, _ = O.calculate_qparams(mn, mx, O.FP4, 4, True, gs)           # fp32 values of e4m3 codes
        codes = s.to(torch.float8_e4m3fn).view(torch.uint8).to(torch.int16)
        codes = (codes + torch.randint(-2, 3, codes.shape, generator=gen, dtype=torch.int16)).clamp(1, 0x7e).to(torch.uint8)
        s = codes.view(torch.float8_e4m3fn).to(torch.bfloat16)
        s.view(-1)[0] = 0.0
        if s.numel() > 2:
            s.view(-1)[1] = 0.3          # not an e4m3 value
        q_o = O.quantize(w, s, torch.zeros(s.shape, dtype=torch.float8_e4m3fn), geom, O.FP4, 4, gs)
        packs.append(q_o[:, 0::2] | (q_o[:, 1::2] << 4))
        scales.append(s)
        # CT quantize (grid values in bf16, -0.0 kept) and fake_quantize with the same foreign scales
        vals = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])[(q_o & 7).long()] * torch.where((q_o & 8) > 0, -1.0, 1.0)
        zfp8 = torch.zeros(s.shape, dtype=torch.float8_e4m3fn)
        q = ops.quantize(w.cuda(), s.cuda(), zfp8.cuda(), args, global_scale=gs.cuda())
        assert_bits_equal(q, vals.to(torch.bfloat16), "nvfp4 quantize values")
        fq = ops.fake_quantize(w.cuda(), s.cuda(), zfp8.cuda(), args, global_scale=gs.cuda())
        assert_bits_equal(fq, O.fake_quantize(w, s, zfp8, geom, O.FP4, 4, gs), "nvfp4 fake_quantize")
    got = ops.quantize_pack(torch.stack(ws).cuda(), torch.stack(scales).cuda(), None, args, global_scale=gs.cuda())
    for i in range(2):
        assert_bits_equal(got[i], packs[i], f"nvfp4[{i}]")


def test_flat_pack_unpack_fast_paths():
    """pack_to_int32 / unpack_from_int32 (4 bit, along the columns) and pack_fp4_to_uint8 / unpack_fp4_from_uint8 (bf16) on shapes that
    take the flat vectorised kernels, including int8 codes outside [-8, 7] (the reference sums the shifted bytes un-masked) and NVFP4 inputs that
    are not e2m1 grid points (first-minimum search in bf16), against the oracle."""
    from quantizers_b200 import ops

    g = torch.Generator().manual_seed(5)
    for R, C in [(16, 24), (37, 768), (128, 2560), (4, 8)]:
        v = torch.randint(-8, 8, (R, C), generator=g, dtype=torch.int8)
        p = ops.pack_to_int32(v.cuda(), 4)
        assert_bits_equal(p, O.pack_to_int32(v, 4), f"pack {R}x{C}")
        assert torch.equal(ops.unpack_from_int32(p, 4, v.shape).cpu(), v)
        assert_bits_equal(ops.unpack_from_int32(p, 4, v.shape), O.unpack_from_int32(O.pack_to_int32(v, 4), 4, (R, C)), f"unpack {R}x{C}")
        wild = torch.randint(-128, 128, (R, C), generator=g, dtype=torch.int8)
        assert_bits_equal(ops.pack_to_int32(wild.cuda(), 4), O.pack_to_int32(wild, 4), f"pack wild {R}x{C}")
        grid = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
        x = (grid[torch.randint(0, 8, (R, C), generator=g)] * torch.where(torch.rand(R, C, generator=g) < 0.5, -1.0, 1.0)).to(torch.bfloat16)
        pk = ops.pack_fp4_to_uint8(x.cuda())
        assert_bits_equal(pk, O.pack_fp4_to_uint8(x), f"pack_fp4 {R}x{C}")
        assert_bits_equal(ops.unpack_fp4_from_uint8(pk, R, C, torch.bfloat16), x, f"unpack_fp4 {R}x{C}")
        off = (torch.randn(R, C, generator=g) * 3).to(torch.bfloat16)   # off-grid, some beyond +-6
        off[0, 0] = -0.0
        assert_bits_equal(ops.pack_fp4_to_uint8(off.cuda()), O.pack_fp4_to_uint8(off), f"pack_fp4 off-grid {R}x{C}")


def test_host_pipeline_refuses_jobs_that_overflow_its_slots():
    """the pipeline's slot buffers are sized from the weight bytes; a CHANNEL scheme on a narrow weight needs more scale
    bytes than that.  The call must fail up front (no device overflow, no truncated D2H), and a normal job must still run after."""
    import ctypes

    from quantizers_b200 import _lib as LB
    from quantizers_b200 import ops
    from quantizers_b200.scheduler import PRESETS

    lib = LB.lib()
    rows, cols = 16384, 8
    h = ctypes.c_void_p()
    LB.check(lib.b200q_pipeline_create(ctypes.byref(h), rows * cols * 2, 0))
    try:
        w = torch.randn(rows, cols).to(torch.bfloat16).pin_memory()
        codes = torch.empty(rows * cols, dtype=torch.uint8).pin_memory()
        scale = torch.empty(rows, dtype=torch.bfloat16).pin_memory()
        sc = ops.scheme_from_args(PRESETS["FP8_CHANNEL"], torch.bfloat16, True)
        rc = lib.b200q_pipeline_compress_host(h, LB.ptr(w), 1, rows, cols, ctypes.byref(sc), LB.ptr(codes), LB.ptr(scale), None, None)
        assert rc != 0 and b"exceed the pipeline slot" in lib.b200q_last_error()
        # a job that fits still works on the same handle, asymmetric INT4 through the zero-point workspace
        w2 = synth_weight(64, 256, torch.bfloat16, 3).pin_memory()
        a = PRESETS["W4A16_ASYM"]
        sc2 = ops.scheme_from_args(a, torch.bfloat16, True)
        c2 = torch.empty((64, 32), dtype=torch.int32).pin_memory()
        s2 = torch.empty((64, 2), dtype=torch.bfloat16).pin_memory()
        z2 = torch.empty((8, 2), dtype=torch.int32).pin_memory()
        LB.check(lib.b200q_pipeline_compress_host(h, LB.ptr(w2), 1, 64, 256, ctypes.byref(sc2), LB.ptr(c2), LB.ptr(s2), LB.ptr(z2), None))
        LB.check(lib.b200q_pipeline_sync(h))
        want = O.compress(w2, "pack-quantized", O.Geom(O.GROUP, 128), 4, False)
        assert_bits_equal(c2, want["weight_packed"], "pipeline codes")
        assert_bits_equal(s2, want["weight_scale"], "pipeline scale")
        assert_bits_equal(z2, want["weight_zero_point"], "pipeline zero points")
    finally:
        lib.b200q_pipeline_destroy(h)
