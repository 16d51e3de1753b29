"""Generate the committed golden fixtures from the LIVE compressed_tensors package (the reference's arithmetic).

Run in the authoring container (compressed_tensors 0.15.0.1 importable):

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden.py

Writes tests/golden/<format>_<dtype>.npz: seeded inputs (bit patterns) + every tensor of the state dict that
``BaseCompressor.compress`` returns (CT:compressors/*/base.py), plus qparams / fake-quant / KAT vectors.
The reference repo itself (/root/reference/tests) holds no numeric vectors for this path.
"""
import os
import sys

os.environ["TORCHDYNAMO_DISABLE"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ct_live as L  # noqa: E402

DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


def bits(t: torch.Tensor) -> np.ndarray:
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.contiguous().view(torch.int16).numpy()
    if t.dtype == torch.float8_e4m3fn:
        return t.contiguous().view(torch.uint8).numpy()
    return t.contiguous().numpy()


def synth(R, C, dtype, seed):
    """SURVEY.md §8d synthetic weight + the edge rows the reference path special-cases."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(R, C, generator=g) * 0.02
    w[:, 3] *= 20  # outlier column
    w[1, :] = -0.0  # exact negative zeros (fp4 sign nibble, fp8 0x00 vs 0x80)
    w[2, : min(C, 128)] = 0.0  # all-zero group -> eps scale, NaN zero-point path
    w[3, :16] = 1e-30  # underflowing scale
    w[4, :] = w[4, :].abs()
    w[5, :] = -w[5, :].abs()
    return w.to(dtype)


def main():
    shapes = {"default": (24, 256), "block": (136, 384)}
    for name in L.FORMATS:
        fmt, args = L.format_args(name)
        for dn, dt in DT.items():
            if dn != "bf16" and name not in ("int4_g128_asym", "int4_g32_sym", "fp8_block", "fp8_channel", "nvfp4"):
                continue
            R, C = shapes["block"] if args.strategy == "block" else shapes["default"]
            w = synth(R, C, dt, seed=1234 + len(name))
            gs = L.global_scale(w) if args.strategy == "tensor_group" else None
            scale, zp = L.weight_qparams(w, args, gs)
            sd = L.compress(w, fmt, args, gs)
            fq = L.fake_quantize(w, scale, zp, args, gs)
            out = {"w": bits(w), "qp_scale": bits(scale), "fq": bits(fq)}
            if zp.dtype == torch.int8:
                out["qp_zp"] = zp.numpy()
            for k, v in sd.items():
                out["sd_" + k] = bits(v)
            np.savez_compressed(os.path.join(HERE, f"{name}_{dn}.npz"), **out)
            print(name, dn, {k: v.shape for k, v in out.items()})

    # known-answer vectors
    from compressed_tensors.compressors.nvfp4.helpers import pack_fp4_to_uint8
    from compressed_tensors.compressors.pack_quantized.helpers import pack_to_int32
    from compressed_tensors.quantization.utils.helpers import generate_gparam

    kat = {}
    kat["pack8_in"] = np.array([[1, 2]], dtype=np.int8)
    kat["pack8_out"] = pack_to_int32(torch.tensor([[1, 2]], dtype=torch.int8), 8).numpy()  # docstring: 33409
    v = torch.arange(-8, 8, dtype=torch.int8).reshape(1, 16).repeat(11, 1)
    kat["pack4_in"] = v.numpy()
    kat["pack4_out"] = pack_to_int32(v, 4).numpy()
    kat["pack4_dim0_out"] = pack_to_int32(v, 4, packed_dim=0).numpy()
    # fp4 edge row: quantize with unit scales, then pack
    args = L.make_args("float", 4, True, "tensor_group", 16)
    row = torch.tensor([[-0.1, 0.1, -0.0, 0.0, -0.25, 0.25, -0.26, 6.0, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, 5.01, -7.0]],
                       dtype=torch.bfloat16)
    from compressed_tensors.quantization.lifecycle.forward import quantize

    q = quantize(row, torch.ones(1, 1, dtype=torch.bfloat16), torch.zeros(1, 1, dtype=torch.float8_e4m3fn), args,
                 global_scale=torch.ones(1))
    kat["fp4_row"] = bits(row)
    kat["fp4_bytes"] = pack_fp4_to_uint8(q).numpy()
    # generate_gparam over a sweep of bf16 / f16 / f32 absmax values (two roundings: reciprocal() * 2688)
    for dn, dt in DT.items():
        a = torch.cat([torch.logspace(-12, 4, 4001), torch.tensor([0.0, 1e-45, 3e38])]).to(dt)
        g = torch.stack([generate_gparam(-x.reshape(1), x.reshape(1)) for x in a]).reshape(-1)
        kat[f"gparam_in_{dn}"] = bits(a)
        kat[f"gparam_out_{dn}"] = g.numpy()
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)
    print("kat", {k: v.shape for k, v in kat.items()})


if __name__ == "__main__":
    main()
