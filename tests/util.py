"""Shared helpers for the parity tests (oracle <-> golden <-> CUDA path)."""
import os

import numpy as np
import torch

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}

# name: (compressor format, qtype, num_bits, symmetric, strategy, group, block)  -- mirrors oracle/ct_live.py
FORMATS = {
    "int4_g128_asym": ("pack-quantized", O.INT, 4, False, O.GROUP, 128, None),
    "int4_g128_sym": ("pack-quantized", O.INT, 4, True, O.GROUP, 128, None),
    "int4_g32_sym": ("pack-quantized", O.INT, 4, True, O.GROUP, 32, None),
    "int4_g32_asym": ("pack-quantized", O.INT, 4, False, O.GROUP, 32, None),
    "int4_channel_sym": ("pack-quantized", O.INT, 4, True, O.CHANNEL, 0, None),
    "int4_channel_asym": ("pack-quantized", O.INT, 4, False, O.CHANNEL, 0, None),
    "int8_g128_sym": ("pack-quantized", O.INT, 8, True, O.GROUP, 128, None),
    "int8_channel_sym": ("pack-quantized", O.INT, 8, True, O.CHANNEL, 0, None),   # W8A8 presets' weights
    "fp8_channel": ("float-quantized", O.FP8, 8, True, O.CHANNEL, 0, None),
    "fp8_g32": ("float-quantized", O.FP8, 8, True, O.GROUP, 32, None),
    "fp8_g128": ("float-quantized", O.FP8, 8, True, O.GROUP, 128, None),
    "fp8_block": ("float-quantized", O.FP8, 8, True, O.BLOCK, 0, (128, 128)),
    "fp8_tensor": ("float-quantized", O.FP8, 8, True, O.TENSOR, 0, None),
    "nvfp4": ("nvfp4-pack-quantized", O.FP4, 4, True, O.GROUP, 16, None),
}


def geom_of(name):
    _, _, _, _, strat, g, blk = FORMATS[name]
    bh, bw = blk or (128, 128)
    return O.Geom(strat, g, bh, bw)


def golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and f != "kat.npz")


def load_golden(fname):
    name, dn = fname[:-4].rsplit("_", 1)
    z = np.load(os.path.join(GOLDEN, fname))
    return name, DT[dn], {k: z[k] for k in z.files}


def from_bits(a: np.ndarray, dtype) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype in (torch.bfloat16, torch.float16) and t.dtype == torch.int16:
        return t.view(dtype)
    return t


def to_bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.view(torch.int16).numpy()
    if t.dtype == torch.float8_e4m3fn:
        return t.view(torch.uint8).numpy()
    return t.numpy()


def assert_bits_equal(got: torch.Tensor, want, what=""):
    g = to_bits(got) if isinstance(got, torch.Tensor) else got
    w = to_bits(want) if isinstance(want, torch.Tensor) else want
    assert g.shape == w.shape, f"{what}: shape {g.shape} != {w.shape}"
    assert g.dtype == w.dtype, f"{what}: dtype {g.dtype} != {w.dtype}"
    bad = int((g != w).sum())
    assert bad == 0, f"{what}: {bad}/{g.size} elements differ"


def synth_weight(R, C, dtype, seed, device="cpu", edge=True):
    """SURVEY.md §8d synthetic weights: randn*0.02, 20x outlier column, zero group, -0.0 row, tiny values."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(R, C, generator=g) * 0.02
    if edge and R >= 6 and C >= 16:
        w[:, 3] *= 20
        w[1, :] = -0.0
        w[2, : min(C, 128)] = 0.0
        w[3, :16] = 1e-30
        w[4, :] = w[4, :].abs()
        w[5, :] = -w[5, :].abs()
    return w.to(dtype).to(device)
