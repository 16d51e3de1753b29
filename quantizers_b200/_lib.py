"""ctypes binding of libb200q.so (include/b200q.h).  No fallback: if the CUDA library is missing or a call
fails, the caller gets an exception -- the product path never silently routes through PyTorch eager or the
CPU oracle."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200q.so")

BF16, F16, F32 = 0, 1, 2
INT, FP8, FP4 = 0, 1, 2
TENSOR, CHANNEL, GROUP, BLOCK = 0, 1, 2, 3

DTYPE_CODE = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}


class B200QError(ValueError):
    """Raised for shape / dtype violations, mirroring the ValueErrors compressed-tensors raises."""


class Scheme(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("dtype", "qtype", "num_bits", "symmetric", "strategy", "group_size", "block_h",
                                       "block_w", "has_zp", "reserved")]


_I = ctypes.c_int
_P = c_void_p
_S = POINTER(Scheme)
_SIGS = {
    "b200q_version": (_I, []),
    "b200q_last_error": (c_char_p, []),
    "b200q_compress_int_packed": (_I, [_P, c_int64, c_int64, c_int64, _S, _P, _P, _P, _P]),
    "b200q_compress_int_workspace": (c_int64, [c_int64, c_int64, c_int64, _S]),
    "b200q_compress_int_packed_ws": (_I, [_P, c_int64, c_int64, c_int64, _S, _P, _P, _P, _P, c_int64, _P]),
    "b200q_compress_fp8": (_I, [_P, c_int64, c_int64, c_int64, _S, _P, _P, _P, _P]),
    "b200q_compress_nvfp4": (_I, [_P, c_int64, c_int64, c_int64, c_int32, c_int32, _P, _P, _P, _P]),
    "b200q_compress_nvfp4_fused": (_I, [_P, c_int64, c_int64, c_int64, c_int32, c_int32, _P, _P, _P, _P, c_int64, _P]),
    "b200q_minmax": (_I, [_P, c_int64, c_int64, c_int64, _S, _P, _P, _P]),
    "b200q_global_scale": (_I, [_P, c_int64, c_int64, c_int32, _P, c_int32, _P, _P]),
    "b200q_calculate_qparams": (_I, [_P, _P, c_int64, _S, _P, _P, _P, _P]),
    "b200q_quantize": (_I, [_P, c_int64, c_int64, _S, _P, _P, _P, _P, _P]),
    "b200q_quantize_pack": (_I, [_P, c_int64, c_int64, c_int64, _S, _P, _P, _P, _P, _P]),
    "b200q_fake_quantize": (_I, [_P, c_int64, c_int64, _S, _P, _P, _P, _P, _P]),
    "b200q_dequantize": (_I, [_P, c_int64, c_int64, _S, _P, _P, _P, _P, _P]),
    "b200q_pack_int32": (_I, [_P, c_int64, c_int64, c_int32, c_int32, _P, _P]),
    "b200q_unpack_int32": (_I, [_P, c_int64, c_int64, c_int32, c_int32, _P, _P]),
    "b200q_pack_fp4": (_I, [_P, c_int64, c_int64, c_int32, _P, _P]),
    "b200q_unpack_fp4": (_I, [_P, c_int64, c_int64, c_int32, _P, _P]),
    "b200q_abs_sum_cols": (_I, [_P, c_int64, c_int64, c_int32, _P, _P]),
    "b200q_wmean_accumulate": (_I, [_P, c_int64, c_int64, c_int32, c_int32, _P, _P]),
    "b200q_awq_scales": (_I, [_P, _P, c_int64, POINTER(c_float), c_int32, c_int32, _P, _P]),
    "b200q_awq_scaled_fake_quantize": (_I, [_P, c_int64, c_int64, _S, _P, _P, _P]),
    "b200q_awq_scaled_fake_quantize_grid": (_I, [_P, c_int64, c_int64, _S, _P, c_int32, _P, c_int64, _P]),
    "b200q_sq_err_accumulate": (_I, [_P, _P, c_int64, c_int32, _P, _P]),
    "b200q_awq_gemm_loss": (_I, [_P, c_int64, c_int64, _P, _P, c_int64, c_int32, _P, _P, c_int64, _P]),
    "b200q_awq_gemm_loss_workspace": (c_int64, [c_int64, c_int64, c_int64, c_int32]),
    "b200q_awq_gemm_loss_pairs": (_I, [_P, _P, c_int64, c_int64, _P, _P, c_int64, c_int32, _P, _P, c_int64, _P]),
    "b200q_qk_norm_rope": (_I, [_P, c_int64, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, _P, c_float, _P]),
    "b200q_attention_workspace": (c_int64, [c_int64, c_int32, c_int32, c_int32]),
    "b200q_attention_core": (_I, [_P, c_int64, c_int32, c_int32, c_int32, c_int32, _P, _P, c_int64, _P]),
    "b200q_awq_gemm_project": (_I, [_P, c_int64, c_int64, _P, c_int64, c_int64, c_int32, _P, _P]),
    "b200q_mse_minmax": (_I, [_P, c_int64, c_int64, c_int64, _S, _P, c_float, c_int32, c_int32, c_float, _P, _P, _P, c_int64, _P]),
    "b200q_compress_nvfp4_workspace": (c_int64, [c_int64, c_int64, c_int64, c_int32]),
    "b200q_awq_gemm_project_grouped": (_I, [_P, c_int64, c_int64, _P, c_int64, c_int64, c_int32, _P, _P, _P]),
    "b200q_moe_combine": (_I, [_P, _P, _P, c_int64, c_int32, c_int64, _P, _P]),
    "b200q_moe_combine_acc": (_I, [_P, _P, _P, c_int64, c_int32, c_int64, _P, _P, _P]),
    "b200q_decompress_int_packed": (_I, [_P, _P, _P, c_int64, c_int64, c_int64, _S, _P, _P]),
    "b200q_decompress_nvfp4": (_I, [_P, _P, _P, c_int64, c_int64, c_int64, c_int32, _P, _P]),
    "b200q_gptq_hessian_accumulate": (_I, [_P, c_int64, c_int64, c_float, c_float, _P, _P]),
    "b200q_pipeline_create": (_I, [POINTER(c_void_p), c_int64, c_int32]),
    "b200q_pipeline_destroy": (_I, [_P]),
    "b200q_pipeline_compress_host": (_I, [_P, _P, c_int64, c_int64, c_int64, _S, _P, _P, _P, _P]),
    "b200q_pipeline_sync": (_I, [_P]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libb200q.so; raise loudly when it has not been built (python -m quantizers_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA extension with `python -m quantizers_b200.build` "
                "(there is no CPU or PyTorch fallback for the quantization hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return list(_SIGS)


def check(rc: int):
    if rc != 0:
        msg = lib().b200q_last_error().decode("utf-8", "replace")
        if rc == -22:
            raise B200QError(msg)
        raise RuntimeError(f"libb200q error {rc}: {msg}")


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "quantizers_b200 ops take CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))


def make_scheme(dtype, qtype, num_bits, symmetric, strategy, group_size=0, block=(128, 128), has_zp=True) -> Scheme:
    if dtype not in DTYPE_CODE:
        raise B200QError(f"unsupported weight dtype {dtype}")
    return Scheme(DTYPE_CODE[dtype], qtype, num_bits, int(bool(symmetric)), strategy, int(group_size or 0), int(block[0]),
                  int(block[1]), int(bool(has_zp)), 0)
