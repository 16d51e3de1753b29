"""ctypes front-end of ``ct_oracle.c`` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Tensors are CPU ``torch`` tensors (torch is only the bf16/fp16 container here; all arithmetic
happens in the C file).  Names mirror the compressed-tensors functions they restate:

  minmax           llmcompressor observers/min_max.py::_get_min_max over flatten_for_calibration
  calculate_qparams  CT:quantization/utils/helpers.py:50-137
  generate_gparam    CT:quantization/utils/helpers.py:309-338
  quantize / fake_quantize / dequantize   CT:quantization/lifecycle/forward.py:37-181
  pack_to_int32 / unpack_from_int32       CT:compressors/pack_quantized/helpers.py:20-161
  pack_fp4_to_uint8 / unpack_fp4_from_uint8   CT:compressors/nvfp4/helpers.py:34-111
  compress         CT:compressors/{pack_quantized,naive_quantized,nvfp4}/base.py ``compress``
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libct_oracle.so")

BF16, F16, F32 = 0, 1, 2
INT, FP8, FP4 = 0, 1, 2
TENSOR, CHANNEL, GROUP, BLOCK = 0, 1, 2, 3

_DT = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}
_TD = {v: k for k, v in _DT.items()}


def build(force: bool = False) -> str:
    """Compile ct_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "ct_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_gparam.restype = ctypes.c_float
        _lib.orc_gparam.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int]
        _lib.orc_compress.restype = ctypes.c_int
        _lib.orc_f32_to_e4m3.restype = ctypes.c_uint8
        _lib.orc_f32_to_e4m3.argtypes = [ctypes.c_float]
        _lib.orc_e4m3_to_f32.restype = ctypes.c_float
        _lib.orc_e4m3_to_f32.argtypes = [ctypes.c_uint8]
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _i64(v):
    return ctypes.c_int64(int(v))


@dataclass
class Geom:
    strategy: int
    group: int = 0
    bh: int = 128
    bw: int = 128

    def qshape(self, R, C):
        out = (ctypes.c_int64 * 2)()
        lib().orc_qparam_shape(self.strategy, _i64(R), _i64(C), _i64(self.group), _i64(self.bh), _i64(self.bw), out)
        if self.strategy == TENSOR:
            return (1,)  # CT:quantization/lifecycle/initialize.py:178
        return int(out[0]), int(out[1])


def _c(t):
    assert t.device.type == "cpu"
    return t.contiguous()


def minmax(w: torch.Tensor, geom: Geom):
    w = _c(w)
    R, C = w.shape
    shp = geom.qshape(R, C)
    mn = torch.empty(shp, dtype=w.dtype)
    mx = torch.empty(shp, dtype=w.dtype)
    lib().orc_minmax(_p(w), _DT[w.dtype], geom.strategy, _i64(R), _i64(C), _i64(geom.group), _i64(geom.bh),
                     _i64(geom.bw), _p(mn), _p(mx))
    return mn, mx


def generate_gparam(mn: float, mx: float, dtype: torch.dtype) -> torch.Tensor:
    return torch.tensor([lib().orc_gparam(float(mn), float(mx), _DT[dtype])], dtype=torch.float32)


def calculate_qparams(mn, mx, qtype, num_bits, symmetric, global_scale=None):
    mn, mx = _c(mn), _c(mx)
    n = mn.numel()
    sdt = torch.float32 if global_scale is not None else mn.dtype
    scale = torch.empty(mn.shape, dtype=sdt)
    zp = torch.zeros(mn.shape, dtype=torch.int8)
    gs = _c(global_scale.float()) if global_scale is not None else None
    lib().orc_qparams(_p(mn), _p(mx), _DT[mn.dtype], _i64(n), qtype, num_bits, int(symmetric), _p(gs), _p(scale), _p(zp))
    return scale, zp


def _quant_args(w, geom, scale, zp, gs):
    w, scale = _c(w), _c(scale)
    if scale.dtype == torch.float8_e4m3fn:
        scale = scale.to(w.dtype)
    zp8 = _c(zp.to(torch.int8)) if zp is not None and zp.dtype in (torch.int8, torch.int32, torch.int64) else None
    has_zp = 1 if zp is not None else 0
    gsf = _c(gs.float()) if gs is not None else None
    R, C = w.shape
    return w, scale, zp8, has_zp, gsf, R, C


def quantize(w, scale, zp, geom: Geom, qtype, num_bits, global_scale=None) -> torch.Tensor:
    """codes: INT -> int8 ; FP8 -> float8_e4m3fn ; FP4 -> uint8 nibble per element (idx | sign<<3)."""
    w, scale, zp8, has_zp, gsf, R, C = _quant_args(w, geom, scale, zp, global_scale)
    out = torch.empty((R, C), dtype=torch.uint8)
    lib().orc_quantize(_p(w), _DT[w.dtype], geom.strategy, _i64(R), _i64(C), _i64(geom.group), _i64(geom.bh),
                       _i64(geom.bw), qtype, num_bits, _p(scale), _DT[scale.dtype], _p(zp8), has_zp, _p(gsf), _p(out))
    if qtype == INT:
        return out.view(torch.int8)
    if qtype == FP8:
        return out.view(torch.float8_e4m3fn)
    return out


def fake_quantize(w, scale, zp, geom: Geom, qtype, num_bits, global_scale=None) -> torch.Tensor:
    w, scale, zp8, has_zp, gsf, R, C = _quant_args(w, geom, scale, zp, global_scale)
    out = torch.empty((R, C), dtype=w.dtype)
    lib().orc_fake_quantize(_p(w), _DT[w.dtype], geom.strategy, _i64(R), _i64(C), _i64(geom.group), _i64(geom.bh),
                            _i64(geom.bw), qtype, num_bits, _p(scale), _DT[scale.dtype], _p(zp8), has_zp, _p(gsf), _p(out))
    return out


def dequantize(codes, scale, zp, geom: Geom, qtype, global_scale=None, out_dtype=None) -> torch.Tensor:
    scale = _c(scale)
    out_dtype = out_dtype or scale.dtype
    R, C = codes.shape
    vals = None
    cbytes = None
    if qtype == FP4:
        vals = _c(codes.float())
    else:
        cbytes = _c(codes.view(torch.uint8))
    zp8 = _c(zp.to(torch.int8)) if zp is not None and qtype == INT else None
    gsf = _c(global_scale.float()) if global_scale is not None else None
    out = torch.empty((R, C), dtype=out_dtype)
    lib().orc_dequantize(_p(cbytes), _p(vals), geom.strategy, _i64(R), _i64(C), _i64(geom.group), _i64(geom.bh),
                         _i64(geom.bw), qtype, _p(scale), _DT[scale.dtype], _p(zp8), 1 if zp is not None else 0,
                         _p(gsf), _p(out), _DT[out_dtype])
    return out


def pack_to_int32(v: torch.Tensor, num_bits: int, packed_dim: int = 1) -> torch.Tensor:
    v = _c(v)
    R, C = v.shape
    pf = 32 // num_bits
    shp = (R, -(-C // pf)) if packed_dim == 1 else (-(-R // pf), C)
    out = torch.empty(shp, dtype=torch.int32)
    lib().orc_pack_int32(_p(v), _i64(R), _i64(C), num_bits, packed_dim, _p(out))
    return out


def unpack_from_int32(p: torch.Tensor, num_bits: int, shape, packed_dim: int = 1) -> torch.Tensor:
    p = _c(p)
    R, C = int(shape[0]), int(shape[1])
    out = torch.empty((R, C), dtype=torch.int8)
    lib().orc_unpack_int32(_p(p), _i64(R), _i64(C), num_bits, packed_dim, _p(out))
    return out


def pack_fp4_to_uint8(x: torch.Tensor) -> torch.Tensor:
    x = _c(x)
    m, n = x.shape
    out = torch.empty((m, n // 2), dtype=torch.uint8)
    lib().orc_pack_fp4(_p(x), _DT[x.dtype], _i64(m), _i64(n), _p(out))
    return out


def unpack_fp4_from_uint8(a: torch.Tensor, m: int, n: int, dtype=torch.bfloat16) -> torch.Tensor:
    a = _c(a)
    out = torch.empty((m, n), dtype=dtype)
    lib().orc_unpack_fp4(_p(a), _i64(m), _i64(n), _p(out), _DT[dtype])
    return out


def compress(w: torch.Tensor, fmt: str, geom: Geom, num_bits: int = 4, symmetric: bool = True, global_scale=None):
    """Observer -> qparams -> quantize -> pack for one weight; returns the CT state-dict entries.

    fmt: "pack-quantized" | "float-quantized" | "nvfp4-pack-quantized"
    """
    w = _c(w)
    R, C = w.shape
    qs = geom.qshape(R, C)
    out = {}
    if fmt == "pack-quantized":
        pf = 32 // num_bits
        packed = torch.empty((R, -(-C // pf)), dtype=torch.int32)
        scale = torch.empty(qs, dtype=w.dtype)
        zpp = torch.empty((-(-qs[0] // pf), qs[-1] if len(qs) > 1 else 1), dtype=torch.int32) if not symmetric else None
        rc = lib().orc_compress(_p(w), _DT[w.dtype], _i64(R), _i64(C), 0, geom.strategy, _i64(geom.group), _i64(geom.bh),
                                _i64(geom.bw), num_bits, int(symmetric), _p(None), _p(packed), _p(scale), _p(zpp), _p(None))
        out = {"weight_packed": packed, "weight_scale": scale, "weight_shape": torch.tensor([R, C])}
        if zpp is not None:
            out["weight_zero_point"] = zpp
    elif fmt == "float-quantized":
        q = torch.empty((R, C), dtype=torch.uint8)
        scale = torch.empty(qs, dtype=w.dtype)
        rc = lib().orc_compress(_p(w), _DT[w.dtype], _i64(R), _i64(C), 1, geom.strategy, _i64(geom.group), _i64(geom.bh),
                                _i64(geom.bw), 8, 1, _p(None), _p(q), _p(scale), _p(None), _p(None))
        out = {"weight": q.view(torch.float8_e4m3fn), "weight_scale": scale}
    elif fmt == "nvfp4-pack-quantized":
        q = torch.empty((R, C // 2), dtype=torch.uint8)
        scale = torch.empty(qs, dtype=torch.uint8)
        gs_out = torch.empty(1, dtype=torch.float32)
        gs_in = _c(global_scale.float()) if global_scale is not None else None
        rc = lib().orc_compress(_p(w), _DT[w.dtype], _i64(R), _i64(C), 2, GROUP, _i64(16), _i64(0), _i64(0), 4, 1,
                                _p(gs_in), _p(q), _p(scale), _p(None), _p(gs_out))
        out = {"weight_packed": q, "weight_scale": scale.view(torch.float8_e4m3fn), "weight_global_scale": gs_out}
    else:
        raise ValueError(fmt)
    if rc != 0:
        raise RuntimeError(f"orc_compress failed: {rc}")
    return out


def abs_sum_cols(x: torch.Tensor) -> torch.Tensor:
    x = _c(x)
    T, K = x.shape
    out = torch.empty(K, dtype=torch.float64)
    lib().orc_abs_sum_cols(_p(x), _DT[x.dtype], _i64(T), _i64(K), _p(out))
    return out


def w_mean(weights, group: int) -> torch.Tensor:
    """llmcompressor AWQModifier._compute_layer_means (restated): mean over all balance-layer rows."""
    K = weights[0].shape[1]
    acc = torch.zeros(K, dtype=torch.float64)
    n = 0
    for w in weights:
        w = _c(w)
        lib().orc_w_mean_group(_p(w), _DT[w.dtype], _i64(w.shape[0]), _i64(K), _i64(group), _p(acc))
        n += w.shape[0]
    return acc / n
