"""CUDA-backed drop-ins for the compressed-tensors callables on the quantization hot path.

Same names, argument meaning and error behaviour as the reference functions:

  quantize / dequantize / fake_quantize      CT:quantization/lifecycle/forward.py:37,77,149
  calculate_qparams / generate_gparam        CT:quantization/utils/helpers.py:50,309
  pack_to_int32 / unpack_from_int32          CT:compressors/pack_quantized/helpers.py:20,92
  pack_fp4_to_uint8 / unpack_fp4_from_uint8  CT:compressors/nvfp4/helpers.py:34,78

plus the fused entry the compressor overrides use (``compress_weight``) and the observer reductions
(``observe_minmax``, ``observe_global_scale``).  Everything runs through the C ABI in include/b200q.h on the
current CUDA stream; tensors must live on the GPU -- there is no CPU path here.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib as L
from ._lib import B200QError

_FP8 = torch.float8_e4m3fn


# ----------------------------------------------------------------------------- QuantizationArgs -> b200q_scheme
def _strategy_code(args) -> int:
    s = getattr(args.strategy, "value", args.strategy)
    if s == "tensor":
        return L.TENSOR
    if s == "channel":
        return L.CHANNEL
    if s in ("group", "tensor_group"):
        return L.GROUP
    if s == "block":
        return L.BLOCK
    raise B200QError(f"quantization strategy {s!r} is not on the weight hot path (token/attn_head are activation-only)")


def _qtype_code(args) -> int:
    t = getattr(args.type, "value", args.type)
    if t == "int":
        if args.num_bits not in (4, 8):
            raise B200QError(f"INT num_bits must be 4 or 8, got {args.num_bits}")
        return L.INT
    if t == "float":
        if args.num_bits == 8:
            return L.FP8
        if args.num_bits == 4:
            return L.FP4
        raise NotImplementedError("Only num_bits in (4, 8) are supported")
    raise B200QError(f"Invalid quantization type {t}")


def scheme_from_args(args, dtype: torch.dtype, has_zp: bool = True) -> L.Scheme:
    block = tuple(args.block_structure) if getattr(args, "block_structure", None) else (128, 128)
    return L.make_scheme(dtype, _qtype_code(args), args.num_bits, args.symmetric, _strategy_code(args),
                         getattr(args, "group_size", None) or 0, block, has_zp)


def _check_g_idx(g_idx):
    if g_idx is None or g_idx.device.type == "meta":
        return
    if bool((g_idx == -1).any()):
        return
    raise NotImplementedError("activation ordering (g_idx) is not on the accelerated path; no reference recipe uses it")


def _as2d(x: torch.Tensor):
    if x.ndim == 1:
        return x.reshape(1, -1), 1
    if x.ndim == 2:
        return x, 1
    return x.reshape(-1, x.shape[-1]), int(x.numel() // (x.shape[-1] * x.shape[-2]))


def _qparam_T(t: Optional[torch.Tensor], dtype) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)  # fp8-typed scales hold e4m3 values: exact in bf16/fp16/fp32
    return t.contiguous()


def _zp_int8(zp: Optional[torch.Tensor], qtype: int) -> Optional[torch.Tensor]:
    if zp is None or qtype != L.INT:
        return None
    return zp.to(torch.int8).contiguous()


def _gs(global_scale: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if global_scale is None:
        return None
    return global_scale.to(torch.float32).reshape(-1)[:1].contiguous()


def _expected_qparams(args, rows, cols) -> int:
    s = _strategy_code(args)
    if s == L.TENSOR:
        return 1
    if s == L.CHANNEL:
        return rows
    if s == L.GROUP:
        return rows * (-(-cols // args.group_size))
    bh, bw = args.block_structure
    return (-(-rows // bh)) * (-(-cols // bw))


def _elementwise(op: str, x, scale, zero_point, args, global_scale, out_dtype_codes=None):
    L.require_cuda(x, scale, zero_point, global_scale)
    if x.ndim != 2:
        raise B200QError(f"expected a 2-D tensor, got shape {tuple(x.shape)}")
    rows, cols = x.shape
    strat = _strategy_code(args)
    if strat == L.GROUP and cols >= args.group_size and cols % args.group_size != 0:
        raise B200QError(f"tensor column shape must be divisble by the given group_size {args.group_size} but got {cols}")
    return rows, cols


@torch.no_grad()
def quantize(x, scale, zero_point, args, dtype=None, g_idx=None, global_scale=None):
    """CT quantize(): codes of ``x`` under caller-supplied qparams.  INT -> integers, FP8 -> e4m3, FP4 -> e2m1 grid
    values held in ``x.dtype``; cast to ``dtype`` when given."""
    _check_g_idx(g_idx)
    rows, cols = _elementwise("quantize", x, scale, zero_point, args, global_scale)
    x = x.contiguous()
    qt = _qtype_code(args)
    sc = scheme_from_args(args, x.dtype, has_zp=zero_point is not None)
    if rows * cols == 0:
        return torch.empty_like(x, dtype=dtype or x.dtype)
    scale_t = _qparam_T(scale, x.dtype)
    if scale_t.numel() != _expected_qparams(args, rows, cols):
        raise B200QError(f"scale has {scale_t.numel()} entries, expected {_expected_qparams(args, rows, cols)}")
    zp8 = _zp_int8(zero_point, qt)
    gs = _gs(global_scale)
    if qt == L.FP4:
        if gs is None:
            gs = torch.ones(1, dtype=torch.float32, device=x.device)
        out = torch.empty_like(x)
    else:
        out = torch.empty((rows, cols), dtype=torch.int8 if qt == L.INT else torch.uint8, device=x.device)
    L.check(L.lib().b200q_quantize(L.ptr(x), rows, cols, ctypes.byref(sc), L.ptr(scale_t), L.ptr(zp8), L.ptr(gs), L.ptr(out),
                                   L.stream_ptr(x.device)))
    if qt == L.FP8:
        out = out.view(_FP8)
    if dtype is None:
        return out if qt == L.FP4 else out.to(x.dtype)
    return out if out.dtype == dtype else out.to(dtype)


@torch.no_grad()
def fake_quantize(x, scale, zero_point, args, g_idx=None, global_scale=None):
    """CT fake_quantize(): quantize then dequantize in one pass, result in ``x.dtype``."""
    _check_g_idx(g_idx)
    rows, cols = _elementwise("fake_quantize", x, scale, zero_point, args, global_scale)
    x = x.contiguous()
    qt = _qtype_code(args)
    if rows * cols == 0:
        return x.clone()
    sc = scheme_from_args(args, x.dtype, has_zp=zero_point is not None)
    scale_t = _qparam_T(scale, x.dtype)
    if scale_t.numel() != _expected_qparams(args, rows, cols):
        raise B200QError(f"scale has {scale_t.numel()} entries, expected {_expected_qparams(args, rows, cols)}")
    gs = _gs(global_scale)
    if qt == L.FP4 and gs is None:
        gs = torch.ones(1, dtype=torch.float32, device=x.device)
    out = torch.empty_like(x)
    L.check(L.lib().b200q_fake_quantize(L.ptr(x), rows, cols, ctypes.byref(sc), L.ptr(scale_t), L.ptr(_zp_int8(zero_point, qt)),
                                        L.ptr(gs), L.ptr(out), L.stream_ptr(x.device)))
    return out


class _InferredArgs:
    """dequantize(args=None): strategy inferred from the scale shape (CT forward.py:99-130)."""

    def __init__(self, strategy, group_size=None, block_structure=None):
        self.strategy, self.group_size, self.block_structure = strategy, group_size, block_structure
        self.symmetric = True


@torch.no_grad()
def dequantize(x_q, scale, zero_point=None, args=None, dtype=None, g_idx=None, global_scale=None):
    """CT dequantize(): (x_q - zp) * scale in the scale dtype, cast to ``dtype``."""
    _check_g_idx(g_idx)
    L.require_cuda(x_q, scale, zero_point, global_scale)
    if x_q.ndim != 2:
        raise B200QError(f"expected a 2-D tensor, got shape {tuple(x_q.shape)}")
    rows, cols = x_q.shape
    if args is None:
        if scale.ndim in (0, 1):
            args = _InferredArgs("tensor")
        elif scale.ndim == 2:
            if scale.shape[1] == 1:
                args = _InferredArgs("channel")
            elif scale.shape[0] == 1 or scale.shape[0] == rows:
                args = _InferredArgs("group", group_size=int(cols / scale.shape[1]))
            else:
                args = _InferredArgs("block", block_structure=[rows // scale.shape[0], cols // scale.shape[1]])
        else:
            raise B200QError(f"Could not infer a quantization strategy from scale with {scale.ndim} dimmensions. "
                             "Expected 0 or 2 dimmensions.")
    sdt = scale.dtype if scale.dtype in L.DTYPE_CODE else (dtype or torch.bfloat16)
    out_dtype = dtype or sdt
    if x_q.dtype == torch.int8:
        qt = L.INT
        codes = x_q.contiguous()
    elif x_q.dtype in (_FP8, torch.uint8):
        qt = L.FP8
        codes = x_q.contiguous().view(torch.uint8)
    else:
        qt = L.FP4  # e2m1 grid values (or any already-decoded code) held in a float dtype
        codes = x_q.to(sdt).contiguous()
    strat = _strategy_code(args)
    block = tuple(args.block_structure) if getattr(args, "block_structure", None) else (128, 128)
    sc = L.make_scheme(sdt, qt, 8 if qt != L.FP4 else 4, True, strat, getattr(args, "group_size", None) or 0, block,
                       has_zp=zero_point is not None)
    scale_t = _qparam_T(scale.reshape(-1) if scale.ndim == 0 else scale, sdt)
    if strat == L.GROUP and scale.ndim == 2 and scale.shape[0] == 1 and rows != 1:
        scale_t = scale_t.expand(rows, -1).contiguous()
    out = torch.empty((rows, cols), dtype=sdt, device=x_q.device)
    if rows * cols:
        L.check(L.lib().b200q_dequantize(L.ptr(codes), rows, cols, ctypes.byref(sc), L.ptr(scale_t), L.ptr(_zp_int8(zero_point, qt)),
                                         L.ptr(_gs(global_scale)), L.ptr(out), L.stream_ptr(x_q.device)))
    return out if out_dtype == sdt else out.to(out_dtype)


@torch.no_grad()
def calculate_qparams(min_vals, max_vals, quantization_args, global_scale=None):
    """CT calculate_qparams(): (scales, zero_points).  Scales come back in the min/max dtype (fp32 when a global
    scale is given), zero-points in ``args.zp_dtype`` (int8 for INT, fp8 zeros otherwise)."""
    L.require_cuda(min_vals, max_vals, global_scale)
    args = quantization_args
    qt = _qtype_code(args)
    if not args.symmetric and qt == L.FP4:
        raise NotImplementedError("Asymmetric Quantization is not supported for FP4")
    dt = min_vals.dtype
    sc = scheme_from_args(args, dt)
    mn, mx = min_vals.contiguous(), max_vals.contiguous()
    gs = _gs(global_scale)
    scale = torch.empty(mn.shape, dtype=torch.float32 if gs is not None else dt, device=mn.device)
    zp = torch.zeros(mn.shape, dtype=torch.int8, device=mn.device)
    L.check(L.lib().b200q_calculate_qparams(L.ptr(mn), L.ptr(mx), mn.numel(), ctypes.byref(sc), L.ptr(gs), L.ptr(scale), L.ptr(zp),
                                            L.stream_ptr(mn.device)))
    zp_dtype = getattr(args, "zp_dtype", None) or torch.int8
    if zp_dtype != torch.int8:
        zp = torch.zeros(mn.shape, dtype=zp_dtype, device=mn.device)
    if scale.ndim == 0:
        scale, zp = scale.reshape(1), zp.reshape(1)
    return scale, zp


@torch.no_grad()
def generate_gparam(updated_min_val, updated_max_val, scale_data=None, quant_data=None, dtype=torch.float32):
    """CT generate_gparam(): 448*6 / absmax with the reference's two roundings; NaN/Inf -> 1.  fp32 [1]."""
    L.require_cuda(updated_min_val, updated_max_val)
    mm = torch.stack([updated_min_val.reshape(-1)[0], updated_max_val.reshape(-1)[0]]).contiguous()
    return observe_global_scale(mm).to(dtype)


@torch.no_grad()
def observe_global_scale(x: torch.Tensor, state: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LLMC Observer.get_global_scale(x): per-tensor min/max -> generate_gparam -> fp32 [1].

    ``state`` (fp32 [2] = running {min, max}, start at {+inf, -inf}) gives the static_minmax semantics used for
    NVFP4 input_global_scale: min/max accumulate across calibration batches and can be all-reduced (MIN/MAX)
    across token shards before the final call."""
    L.require_cuda(x, state)
    x = x.contiguous()
    running = state is not None
    if state is None:
        state = torch.empty(2, dtype=torch.float32, device=x.device)
    gs = torch.empty(1, dtype=torch.float32, device=x.device)
    L.check(L.lib().b200q_global_scale(L.ptr(x), 1, x.numel(), L.DTYPE_CODE[x.dtype], L.ptr(state), int(running), L.ptr(gs),
                                       L.stream_ptr(x.device)))
    return gs


@torch.no_grad()
def observe_minmax(weight: torch.Tensor, args):
    """memoryless_minmax Observer statistics: (min, max) per quantization chunk, in the weight dtype, shaped like
    the qparams (flatten_for_calibration + amin/amax over dims (0,-1))."""
    L.require_cuda(weight)
    w = weight.contiguous()
    strat = _strategy_code(args)
    if strat == L.TENSOR:
        state = torch.empty(2, dtype=torch.float32, device=w.device)
        L.check(L.lib().b200q_global_scale(L.ptr(w), 1, w.numel(), L.DTYPE_CODE[w.dtype], L.ptr(state), 0, None,
                                           L.stream_ptr(w.device)))
        return state[0:1].to(w.dtype), state[1:2].to(w.dtype)
    w2, batch = _as2d(w)
    rows, cols = w.shape[-2], w.shape[-1]
    sc = scheme_from_args(args, w.dtype)
    if strat == L.CHANNEL:
        shp = (*w.shape[:-1], 1)
    elif strat == L.GROUP:
        if cols % args.group_size != 0:
            raise B200QError(f"tensor column shape must be divisble by the given group_size {args.group_size} but got {cols}")
        shp = (*w.shape[:-1], cols // args.group_size)
    else:
        bh, bw = args.block_structure
        shp = (*w.shape[:-2], -(-rows // bh), -(-cols // bw))
    mn = torch.empty(shp, dtype=w.dtype, device=w.device)
    mx = torch.empty(shp, dtype=w.dtype, device=w.device)
    L.check(L.lib().b200q_minmax(L.ptr(w), batch, rows, cols, ctypes.byref(sc), L.ptr(mn), L.ptr(mx), L.stream_ptr(w.device)))
    return mn, mx


@torch.no_grad()
def observe_mse_minmax(weight: torch.Tensor, args, global_scale: Optional[torch.Tensor] = None, maxshrink: float = 0.2,
                       patience: int = 5, grid: int = 100, norm: float = 2.4):
    """LLMC ``mse`` observer statistics (observers/mse.py): per quantization chunk the shrunk (min, max) range
    ``p * (min, max)``, ``p = 1 - i/grid`` for ``i < int(maxshrink * grid)``, with the smallest ``sum |fake_quantize(x) - x|^norm``;
    tensor-wide early stop after ``patience`` steps without any improvement.  GROUP / TENSOR_GROUP / CHANNEL strategies."""
    L.require_cuda(weight, global_scale)
    w = weight.contiguous()
    strat = _strategy_code(args)
    if strat not in (L.GROUP, L.CHANNEL):
        raise NotImplementedError("mse observer: GROUP / TENSOR_GROUP / CHANNEL strategies only")
    w2, batch = _as2d(w)
    rows, cols = w.shape[-2], w.shape[-1]
    sc = scheme_from_args(args, w.dtype)
    if strat == L.CHANNEL:
        shp = (*w.shape[:-1], 1)
    else:
        if cols % args.group_size != 0:
            raise B200QError(f"tensor column shape must be divisble by the given group_size {args.group_size} but got {cols}")
        shp = (*w.shape[:-1], cols // args.group_size)
    mn = torch.empty(shp, dtype=w.dtype, device=w.device)
    mx = torch.empty(shp, dtype=w.dtype, device=w.device)
    ws = torch.empty(mn.numel() + batch, dtype=torch.int32, device=w.device)
    L.check(L.lib().b200q_mse_minmax(L.ptr(w), batch, rows, cols, ctypes.byref(sc), L.ptr(_gs(global_scale)), float(maxshrink), int(patience),
                                     int(grid), float(norm), L.ptr(mn), L.ptr(mx), L.ptr(ws), ws.numel() * 4, L.stream_ptr(w.device)))
    return mn, mx


# ----------------------------------------------------------------------------- fused decompress
@torch.no_grad()
def decompress_int_packed(packed: torch.Tensor, scale: torch.Tensor, zp_packed: Optional[torch.Tensor], shape, args) -> torch.Tensor:
    """pack-quantized ``Compressor.decompress`` in one pass: ``weight_packed`` int32 + ``weight_scale`` (+ row-packed
    ``weight_zero_point``) -> weight ``[*lead, rows, cols]`` in the scale dtype (unpack_from_int32 + dequantize fused)."""
    L.require_cuda(packed, scale, zp_packed)
    rows, cols = int(shape[-2]), int(shape[-1])
    lead = tuple(packed.shape[:-2])
    batch = 1
    for d in lead:
        batch *= int(d)
    if scale.dtype not in L.DTYPE_CODE:
        raise B200QError(f"unsupported scale dtype {scale.dtype}")
    sc = scheme_from_args(args, scale.dtype, zp_packed is not None)
    out = torch.empty(lead + (rows, cols), dtype=scale.dtype, device=packed.device)
    L.check(L.lib().b200q_decompress_int_packed(L.ptr(packed.contiguous()), L.ptr(scale.contiguous()),
                                                L.ptr(zp_packed.contiguous() if zp_packed is not None else None), batch, rows, cols,
                                                ctypes.byref(sc), L.ptr(out), L.stream_ptr(packed.device)))
    return out


@torch.no_grad()
def decompress_nvfp4(packed: torch.Tensor, scale: torch.Tensor, global_scale: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """nvfp4-pack-quantized ``decompress`` in one pass: u8 code pairs + e4m3 group scales + fp32 global scale -> weight."""
    L.require_cuda(packed, scale, global_scale)
    rows, half = int(packed.shape[-2]), int(packed.shape[-1])
    lead = tuple(packed.shape[:-2])
    batch = 1
    for d in lead:
        batch *= int(d)
    gs = global_scale.to(torch.float32).reshape(-1).contiguous()
    if gs.numel() == 1 and batch > 1:
        gs = gs.expand(batch).contiguous()
    if gs.numel() != batch:
        raise B200QError(f"global_scale must have one entry per stacked weight, got {gs.numel()} for {batch}")
    out = torch.empty(lead + (rows, half * 2), dtype=dtype, device=packed.device)
    L.check(L.lib().b200q_decompress_nvfp4(L.ptr(packed.contiguous()), L.ptr(scale.contiguous().view(torch.uint8)), L.ptr(gs), batch, rows, half * 2,
                                           L.DTYPE_CODE[dtype], L.ptr(out), L.stream_ptr(packed.device)))
    return out


# ----------------------------------------------------------------------------- pack / unpack
@torch.no_grad()
def pack_to_int32(value: torch.Tensor, num_bits: int, packed_dim: int = 1) -> torch.Tensor:
    if value.dtype is not torch.int8:
        raise ValueError("Tensor must be quantized to torch.int8 before packing")
    if num_bits > 8:
        raise ValueError("Packing is only supported for less than 8 bits")
    if num_bits < 1:
        raise ValueError(f"num_bits must be at least 1, got {num_bits}")
    if value.ndim > 2:
        return torch.stack([pack_to_int32(value[i], num_bits, packed_dim) for i in range(value.shape[0])])
    L.require_cuda(value)
    value = value.contiguous()
    rows, cols = value.shape
    pf = 32 // num_bits
    shape = (rows, -(-cols // pf)) if packed_dim == 1 else (-(-rows // pf), cols)
    out = torch.empty(shape, dtype=torch.int32, device=value.device)
    L.check(L.lib().b200q_pack_int32(L.ptr(value), rows, cols, num_bits, packed_dim, L.ptr(out), L.stream_ptr(value.device)))
    return out


@torch.no_grad()
def unpack_from_int32(value: torch.Tensor, num_bits: int, shape, packed_dim: int = 1) -> torch.Tensor:
    if value.dtype is not torch.int32:
        raise ValueError(f"Expected {torch.int32} but got {value.dtype}, Aborting unpack.")
    if num_bits > 8:
        raise ValueError("Unpacking is only supported for less than 8 bits")
    if value.ndim > 2:
        return torch.stack([unpack_from_int32(value[i], num_bits, shape[1:], packed_dim) for i in range(value.shape[0])])
    L.require_cuda(value)
    value = value.contiguous()
    rows, cols = int(shape[0]), int(shape[1])
    out = torch.empty((rows, cols), dtype=torch.int8, device=value.device)
    L.check(L.lib().b200q_unpack_int32(L.ptr(value), rows, cols, num_bits, packed_dim, L.ptr(out), L.stream_ptr(value.device)))
    return out


@torch.no_grad()
def pack_fp4_to_uint8(x: torch.Tensor) -> torch.Tensor:
    m, n = x.shape
    if n % 2 != 0:
        raise ValueError("tensor must have an even number of columns for nvfp4 compression")
    L.require_cuda(x)
    x = x.contiguous()
    out = torch.empty((m, n // 2), dtype=torch.uint8, device=x.device)
    L.check(L.lib().b200q_pack_fp4(L.ptr(x), m, n, L.DTYPE_CODE[x.dtype], L.ptr(out), L.stream_ptr(x.device)))
    return out


@torch.no_grad()
def unpack_fp4_from_uint8(a: torch.Tensor, m: int, n: int, dtype: Optional[torch.dtype] = torch.bfloat16) -> torch.Tensor:
    assert a.dtype == torch.uint8
    L.require_cuda(a)
    a = a.contiguous()
    out = torch.empty((m, n), dtype=dtype, device=a.device)
    L.check(L.lib().b200q_unpack_fp4(L.ptr(a), m, n, L.DTYPE_CODE[dtype], L.ptr(out), L.stream_ptr(a.device)))
    return out


# ----------------------------------------------------------------------------- fused compress
def _out_buf(out: Optional[dict], key: str, shape, dtype, dev) -> torch.Tensor:
    """``out[key]`` when the caller owns the buffer (checked), else a fresh allocation."""
    if out is not None and key in out:
        t = out[key]
        if t.dtype == _FP8 and dtype == torch.uint8:
            t = t.view(torch.uint8)
        if tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev or not t.is_contiguous():
            raise B200QError(f"out[{key!r}] must be a contiguous {dtype} tensor of shape {tuple(shape)} on {dev}, "
                             f"got {t.dtype} {tuple(t.shape)} on {t.device}")
        return t
    return torch.empty(shape, dtype=dtype, device=dev)


def compress_outputs(weight_shape, args, dtype=torch.bfloat16, device="cuda", fuse_span: int = 1) -> dict:
    """Caller-owned output (and workspace) buffers for :func:`compress_weight` on a weight / stack of ``weight_shape``: allocate
    once, pass as ``out=`` for every weight of that shape class -- the call then allocates nothing and never touches the host."""
    lead, rows, cols = tuple(weight_shape[:-2]), int(weight_shape[-2]), int(weight_shape[-1])
    batch = 1
    for d in lead:
        batch *= int(d)
    dev = torch.device(device)
    qt, strat = _qtype_code(args), _strategy_code(args)
    g = getattr(args, "group_size", None) or 0
    if strat == L.GROUP:
        qshape = (rows, cols // g)
    elif strat == L.CHANNEL:
        qshape = (rows, 1)
    elif strat == L.BLOCK:
        bh, bw = args.block_structure
        qshape = (-(-rows // bh), -(-cols // bw))
    else:
        qshape = (1,)
    if qt == L.INT:
        pf = 32 // args.num_bits
        out = {"weight_packed": torch.empty(lead + (rows, -(-cols // pf)), dtype=torch.int32, device=dev),
               "weight_scale": torch.empty(lead + qshape, dtype=dtype, device=dev)}
        if not args.symmetric:
            out["weight_zero_point"] = torch.empty(lead + (-(-qshape[0] // pf), qshape[1]), dtype=torch.int32, device=dev)
            if strat == L.GROUP:
                out["_workspace"] = torch.empty(max(batch * rows * qshape[1], 1), dtype=torch.int8, device=dev)
        return out
    if qt == L.FP8:
        out = {"weight": torch.empty(lead + (rows, cols), dtype=torch.uint8, device=dev).view(_FP8),
               "weight_scale": torch.empty(lead + qshape, dtype=dtype, device=dev)}
        if strat == L.TENSOR:
            out["_workspace"] = torch.empty(max(batch, 1), dtype=torch.float32, device=dev)
        return out
    out = {"weight_packed": torch.empty(lead + (rows, cols // 2), dtype=torch.uint8, device=dev),
           "weight_scale": torch.empty(lead + (rows, cols // 16), dtype=torch.uint8, device=dev).view(_FP8),
           "weight_global_scale": torch.empty(lead + (1,), dtype=torch.float32, device=dev)}
    nws = max(int(L.lib().b200q_compress_nvfp4_workspace(batch, rows, cols, int(fuse_span))) // 4, 2)
    out["_workspace"] = torch.empty(nws, dtype=torch.int32, device=dev)
    return out


@torch.no_grad()
def compress_weight(weight: torch.Tensor, args, global_scale: Optional[torch.Tensor] = None, has_zp: bool = True,
                    fuse_span: int = 1, out: Optional[dict] = None) -> dict:
    """Fused observer -> qparams -> quantize -> pack of one weight ``[rows, cols]`` or a stack ``[E, rows, cols]``.

    Returns the tensors ``Compressor.compress`` puts in the state dict:
      INT  : weight_packed int32, weight_scale T, weight_shape int64[2] on the CPU (CT:compressors/pack_quantized/base.py:68),
             (+ weight_zero_point int32 when asymmetric)
      FP8  : weight e4m3, weight_scale T
      FP4  : weight_packed uint8, weight_scale e4m3, weight_global_scale fp32 [1] (per stacked weight: [E, 1]);
             ``fuse_span`` consecutive stacked weights share min(global_scale) (gate/up siblings, LLMC
             update_fused_layer_weight_global_scales) when the global scale is computed here
    ``out``: buffers from :func:`compress_outputs` (same shape class) -- the results are written there, nothing is allocated and
    no host synchronisation happens; without it every tensor is a fresh allocation.
    """
    L.require_cuda(weight, global_scale)
    if weight.ndim not in (2, 3):
        raise B200QError(f"expected a [rows, cols] weight or an [experts, rows, cols] stack, got {tuple(weight.shape)}")
    w = weight.contiguous()
    batch = 1 if w.ndim == 2 else w.shape[0]
    rows, cols = w.shape[-2], w.shape[-1]
    lead = tuple(w.shape[:-2])
    qt = _qtype_code(args)
    strat = _strategy_code(args)
    dev = w.device
    st = L.stream_ptr(dev)
    sc = scheme_from_args(args, w.dtype, has_zp)
    lib = L.lib()
    if strat == L.GROUP and cols % args.group_size != 0:
        raise B200QError(f"tensor column shape must be divisble by the given group_size {args.group_size} but got {cols}")
    if qt == L.INT:
        pf = 32 // args.num_bits
        if strat == L.GROUP:
            qshape = (rows, cols // args.group_size)
        elif strat == L.CHANNEL:
            qshape = (rows, 1)
        else:
            raise B200QError("pack-quantized supports group and channel strategies")
        packed = _out_buf(out, "weight_packed", lead + (rows, -(-cols // pf)), torch.int32, dev)
        scale = _out_buf(out, "weight_scale", lead + qshape, w.dtype, dev)
        zpp = None if args.symmetric else _out_buf(out, "weight_zero_point", lead + (-(-qshape[0] // pf), qshape[1]), torch.int32, dev)
        ws = None
        if zpp is not None and strat == L.GROUP:
            # int8 zero points land here with plain stores and are row-packed by a second small kernel (no atomics, no memset)
            ws = _out_buf(out, "_workspace", (max(batch * rows * qshape[1], 1),), torch.int8, dev)
        L.check(lib.b200q_compress_int_packed_ws(L.ptr(w), batch, rows, cols, ctypes.byref(sc), L.ptr(packed), L.ptr(scale), L.ptr(zpp),
                                                 L.ptr(ws), 0 if ws is None else ws.numel(), st))
        res = {"weight_packed": packed, "weight_scale": scale, "weight_shape": torch.tensor([rows, cols])}
        if zpp is not None:
            res["weight_zero_point"] = zpp
        return res
    if qt == L.FP8:
        if strat == L.GROUP:
            qshape = (rows, cols // args.group_size)
        elif strat == L.CHANNEL:
            qshape = (rows, 1)
        elif strat == L.BLOCK:
            bh, bw = args.block_structure
            qshape = (-(-rows // bh), -(-cols // bw))
        else:
            qshape = (1,)
        q = _out_buf(out, "weight", lead + (rows, cols), torch.uint8, dev)
        scale = _out_buf(out, "weight_scale", lead + qshape, w.dtype, dev)
        ws = _out_buf(out, "_workspace", (max(batch, 1),), torch.float32, dev) if strat == L.TENSOR else None
        L.check(lib.b200q_compress_fp8(L.ptr(w), batch, rows, cols, ctypes.byref(sc), L.ptr(q), L.ptr(scale), L.ptr(ws), st))
        return {"weight": q.view(_FP8), "weight_scale": scale}
    # NVFP4
    if cols % 16 != 0:
        raise B200QError(f"tensor column shape must be divisble by the given group_size 16 but got {cols}")
    packed = _out_buf(out, "weight_packed", lead + (rows, cols // 2), torch.uint8, dev)
    scale = _out_buf(out, "weight_scale", lead + (rows, cols // 16), torch.uint8, dev)
    if global_scale is None:
        # NVFP4 siblings stacked next to each other (gate/up of one expert: fuse_span = 2) share min(global_scale)
        gs = _out_buf(out, "weight_global_scale", lead + (1,), torch.float32, dev)
        nws = max(int(lib.b200q_compress_nvfp4_workspace(batch, rows, cols, int(fuse_span))) // 4, 2)
        ws = _out_buf(out, "_workspace", (nws,), torch.int32, dev)
        L.check(lib.b200q_compress_nvfp4_fused(L.ptr(w), batch, rows, cols, L.DTYPE_CODE[w.dtype], int(fuse_span), L.ptr(gs), L.ptr(packed),
                                               L.ptr(scale), L.ptr(ws), ws.numel() * 4, st))
        return {"weight_packed": packed, "weight_scale": scale.view(_FP8), "weight_global_scale": gs.reshape(lead + (1,))}
    else:
        gs = global_scale.to(torch.float32).reshape(-1)
        if gs.numel() == 1 and batch > 1:
            gs = gs.expand(batch)
        gs = gs.contiguous()
        compute = 0
    L.check(lib.b200q_compress_nvfp4(L.ptr(w), batch, rows, cols, L.DTYPE_CODE[w.dtype], compute, L.ptr(gs), L.ptr(packed), L.ptr(scale), st))
    return {"weight_packed": packed, "weight_scale": scale.view(_FP8), "weight_global_scale": gs.reshape(lead + (1,))}


@torch.no_grad()
def quantize_pack(weight: torch.Tensor, scale, zero_point, args, global_scale=None) -> torch.Tensor:
    """``Compressor.compress``'s arithmetic with the module's existing qparams: quantize + pack in one pass.
    Returns ``weight_packed`` (INT: int32, FP4: uint8) or the e4m3 ``weight`` (FP8)."""
    L.require_cuda(weight, scale, zero_point, global_scale)
    w = weight.contiguous()
    rows, cols = w.shape[-2], w.shape[-1]
    batch = 1 if w.ndim == 2 else w.shape[0]
    lead = tuple(w.shape[:-2])
    qt = _qtype_code(args)
    strat = _strategy_code(args)
    g = getattr(args, "group_size", None) or 0
    fused = strat == L.GROUP and g in (16, 32, 64, 128, 256) and cols % g == 0
    if not fused:
        if w.ndim != 2:
            return torch.stack([quantize_pack(w[i], scale[i], None if zero_point is None else zero_point[i], args, global_scale)
                                for i in range(batch)])
        if qt == L.INT:
            return pack_to_int32(quantize(w, scale, zero_point, args, dtype=torch.int8), args.num_bits)
        if qt == L.FP8:
            return quantize(w, scale, zero_point, args, dtype=_FP8)
        return pack_fp4_to_uint8(quantize(w, scale, zero_point, args, global_scale=global_scale))
    sc = scheme_from_args(args, w.dtype, has_zp=zero_point is not None)
    scale_t = _qparam_T(scale, w.dtype)
    if scale_t.numel() != batch * rows * (cols // g):
        raise B200QError(f"scale has {scale_t.numel()} entries, expected {batch * rows * (cols // g)}")
    gs = _gs(global_scale)
    if qt == L.INT:
        pf = 32 // args.num_bits
        out = torch.empty(lead + (rows, -(-cols // pf)), dtype=torch.int32, device=w.device)
    elif qt == L.FP8:
        out = torch.empty(lead + (rows, cols), dtype=torch.uint8, device=w.device)
    else:
        if gs is None:
            gs = torch.ones(1, dtype=torch.float32, device=w.device)
        out = torch.empty(lead + (rows, cols // 2), dtype=torch.uint8, device=w.device)
    if w.numel():
        L.check(L.lib().b200q_quantize_pack(L.ptr(w), batch, rows, cols, ctypes.byref(sc), L.ptr(scale_t),
                                            L.ptr(_zp_int8(zero_point, qt)), L.ptr(gs), L.ptr(out), L.stream_ptr(w.device)))
    return out.view(_FP8) if qt == L.FP8 else out


@torch.no_grad()
def weight_global_scales(weights: torch.Tensor) -> torch.Tensor:
    """Batched Observer.get_global_scale over a stack ``[units, rows, cols]`` -> fp32 ``[units]``."""
    L.require_cuda(weights)
    w = weights.contiguous()
    batch = 1 if w.ndim == 2 else w.shape[0]
    state = torch.empty(2 * batch, dtype=torch.float32, device=w.device)
    gs = torch.empty(batch, dtype=torch.float32, device=w.device)
    L.check(L.lib().b200q_global_scale(L.ptr(w), batch, w.numel() // batch, L.DTYPE_CODE[w.dtype], L.ptr(state), 0, L.ptr(gs),
                                       L.stream_ptr(w.device)))
    return gs
