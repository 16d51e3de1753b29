"""CUDA-backed compressors registered under the reference's format names.

``ModelCompressor.compress_model`` -> ``compress_module`` -> ``BaseCompressor.get_value_from_registry(format)``
-> ``compress(state_dict, scheme)`` (CT:compressors/base.py:71-87,145-168) is the seam ``save_pretrained(
save_compressed=True)`` goes through in /root/reference/scripts/do_oneshot.py:197.  The classes below keep the
exact contract of the CT compressors they subclass -- same input keys, input dict not modified, same output
keys / dtypes / shapes -- and run the quantize + pack arithmetic in one CUDA pass.  ``decompress`` runs the
unpack + dequantize kernels.  Use ``quantizers_b200.patch.patch()`` to install them.
"""
from __future__ import annotations

import torch
from compressed_tensors.compressors.nvfp4.base import NVFP4PackedCompressor
from compressed_tensors.compressors.naive_quantized.base import (
    FloatQuantizationCompressor,
    IntQuantizationCompressor,
    NaiveQuantizationCompressor,
)
from compressed_tensors.compressors.pack_quantized.base import PACK_ZP_STRATS, PackedQuantizationCompressor

from . import ops

__all__ = ["B200PackedQuantizationCompressor", "B200NVFP4PackedCompressor", "B200NaiveQuantizationCompressor",
           "B200IntQuantizationCompressor", "B200FloatQuantizationCompressor", "REGISTRY_OVERRIDES"]


def _cuda(t):
    return t is not None and t.is_cuda


class B200PackedQuantizationCompressor(PackedQuantizationCompressor):
    """pack-quantized: CT:compressors/pack_quantized/base.py:36-77."""

    @classmethod
    def compress(cls, state_dict, scheme):
        if not _cuda(state_dict.get("weight")):
            raise RuntimeError("B200PackedQuantizationCompressor needs the weight on a CUDA device (no CPU fallback)")
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        scale = state_dict.get("weight_scale")
        zero_point = state_dict.get("weight_zero_point", None)
        ops._check_g_idx(state_dict.get("weight_g_idx", None))
        weights = scheme.weights
        state_dict["weight_packed"] = ops.quantize_pack(weight, scale, zero_point, weights)
        state_dict["weight_shape"] = torch.tensor(weight.shape)
        if not weights.symmetric and weights.strategy in PACK_ZP_STRATS:
            assert zero_point is not None, "Asymmetric quant requires zero-point values"
            state_dict["weight_zero_point"] = ops.pack_to_int32(zero_point, weights.num_bits, packed_dim=0).contiguous()
        return cls._remove_symmetric_zp(state_dict, scheme)

    @classmethod
    def decompress(cls, state_dict, scheme):
        state_dict = state_dict.copy()
        packed = state_dict.pop("weight_packed")
        scale = state_dict.get("weight_scale")
        zero_point = state_dict.get("weight_zero_point", None)
        original_shape = state_dict.get("weight_shape")
        weights = scheme.weights
        asym = not weights.symmetric and weights.strategy in PACK_ZP_STRATS
        if asym:
            assert zero_point is not None, "Asymmetric quant requires zero-point values"
        strat = getattr(weights.strategy, "value", weights.strategy)
        if packed.is_cuda and strat in ("group", "channel") and scale.dtype in (torch.bfloat16, torch.float16, torch.float32):
            # fused unpack + dequantize: one pass, no int8 intermediate
            state_dict["weight"] = ops.decompress_int_packed(packed, scale, zero_point if asym else None, tuple(int(v) for v in original_shape),
                                                             weights)
            if asym:
                state_dict["weight_zero_point"] = ops.unpack_from_int32(zero_point, weights.num_bits, (*original_shape[:-1], scale.shape[-1]),
                                                                        packed_dim=0)
            return state_dict
        if asym:
            zero_point = ops.unpack_from_int32(zero_point, weights.num_bits, (*original_shape[:-1], scale.shape[-1]), packed_dim=0)
            state_dict["weight_zero_point"] = zero_point
        unpacked = ops.unpack_from_int32(packed, weights.num_bits, original_shape)
        state_dict["weight"] = ops.dequantize(x_q=unpacked, scale=scale, zero_point=zero_point)
        return state_dict


class B200NVFP4PackedCompressor(NVFP4PackedCompressor):
    """nvfp4-pack-quantized: CT:compressors/nvfp4/base.py:40-72."""

    @classmethod
    def compress(cls, state_dict, scheme):
        if not _cuda(state_dict.get("weight")):
            raise RuntimeError("B200NVFP4PackedCompressor needs the weight on a CUDA device (no CPU fallback)")
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        scale = state_dict.pop("weight_scale")
        global_scale = state_dict.get("weight_global_scale", None)
        zero_point = state_dict.get("weight_zero_point", None)
        weights = scheme.weights
        state_dict["weight_packed"] = ops.quantize_pack(weight, scale, zero_point, weights, global_scale=global_scale)
        state_dict["weight_scale"] = cls._compress_scale(scale, weights)
        return cls._remove_symmetric_zp(state_dict, scheme)

    @classmethod
    def decompress(cls, state_dict, scheme):
        state_dict = state_dict.copy()
        packed = state_dict.pop("weight_packed")
        scale = state_dict.get("weight_scale")
        global_scale = state_dict.get("weight_global_scale", None)
        m, n = packed.shape
        if packed.is_cuda and global_scale is not None and scale.dtype == torch.float8_e4m3fn:
            state_dict["weight"] = ops.decompress_nvfp4(packed, scale, global_scale, torch.bfloat16)  # fused unpack + dequantize
            state_dict["weight_scale"] = torch.nn.Parameter(scale.to(torch.bfloat16), requires_grad=False)
            return state_dict
        unpacked = ops.unpack_fp4_from_uint8(packed, m, n * 2)
        scale_float = scale.to(unpacked.dtype)
        state_dict["weight"] = ops.dequantize(x_q=unpacked, scale=scale_float, global_scale=global_scale, dtype=unpacked.dtype)
        state_dict["weight_scale"] = torch.nn.Parameter(scale_float, requires_grad=False)
        return state_dict


class _NaiveMixin:
    """naive / int / float-quantized: CT:compressors/naive_quantized/base.py:34-82.  Block padding is implicit:
    the kernels treat ragged 128x128 edges as zero padded and only ever write the original shape."""

    @classmethod
    def compress(cls, state_dict, scheme):
        if not _cuda(state_dict.get("weight")):
            raise RuntimeError(f"{cls.__name__} needs the weight on a CUDA device (no CPU fallback)")
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        scale = state_dict.get("weight_scale")
        zero_point = state_dict.get("weight_zero_point", None)
        ops._check_g_idx(state_dict.get("weight_g_idx", None))
        weights = scheme.weights
        state_dict["weight"] = ops.quantize(weight, scale, zero_point, weights, dtype=weights.pytorch_dtype())
        return cls._remove_symmetric_zp(state_dict, scheme)

    @classmethod
    def decompress(cls, state_dict, scheme):
        state_dict = state_dict.copy()
        weight = state_dict.pop("weight")
        state_dict["weight"] = ops.dequantize(x_q=weight, scale=state_dict.get("weight_scale"),
                                              zero_point=state_dict.get("weight_zero_point", None))
        return state_dict


class B200NaiveQuantizationCompressor(_NaiveMixin, NaiveQuantizationCompressor):
    pass


class B200IntQuantizationCompressor(_NaiveMixin, IntQuantizationCompressor):
    pass


class B200FloatQuantizationCompressor(_NaiveMixin, FloatQuantizationCompressor):
    pass


REGISTRY_OVERRIDES = {
    "pack-quantized": B200PackedQuantizationCompressor,
    "nvfp4-pack-quantized": B200NVFP4PackedCompressor,
    "naive-quantized": B200NaiveQuantizationCompressor,
    "int-quantized": B200IntQuantizationCompressor,
    "float-quantized": B200FloatQuantizationCompressor,
}
