"""Run a model straight from its compressed tensors: a Linear whose weight stays packed and is decoded on each forward call.

SURVEY.md §8f rank 1 (closes the loop after ``oneshot``: the reference prints a sample generation from the quantized model before
saving it, REF:scripts/quantization_multiple_modifiers.py:111-119).  compressed-tensors used to ship this as
``CompressedLinear`` ("the wrapped layer will be decompressed on each forward call", CT:linear/compressed_linear.py:8-12); the
0.15 wheel keeps the class but raises in ``from_linear``, so the behaviour is restated here on the fused decode kernels:

  pack-quantized        weight_packed int32 + weight_scale (+ row-packed weight_zero_point)  -> ``b200q_decompress_int_packed``
  nvfp4-pack-quantized  weight_packed u8 + e4m3 weight_scale + fp32 weight_global_scale      -> ``b200q_decompress_nvfp4``
  float-/naive-/int-quantized  weight (e4m3 / int8) + weight_scale (+ weight_zero_point)     -> ``b200q_dequantize``

The decoded weight equals ``fake_quantize(original weight)`` (SURVEY.md Appendix B identities; bit for bit for FP8 / NVFP4, and
as numbers for integer formats, where a stored code 0 decodes to +0.0 while fake_quantize may carry -0.0 -- live CT behaves the
same way), so the forward of the swapped model is the forward of the fake-quantized model.  The GEMM itself is a plain library
``F.linear``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch

from . import ops
from .recipe import Recipe, resolve_targets

_KEYS = ("weight_packed", "weight", "weight_scale", "weight_zero_point", "weight_global_scale", "weight_shape")


class CompressedLinear(torch.nn.Module):
    """Holds the compressed tensors of one Linear as buffers (same names as in the checkpoint) and decodes them per call."""

    def __init__(self, entries: Dict[str, torch.Tensor], args, in_features: int, out_features: int, bias: Optional[torch.Tensor] = None,
                 dtype: torch.dtype = torch.bfloat16, cache: bool = False):
        super().__init__()
        unknown = set(entries) - set(_KEYS)
        if unknown:
            raise ValueError(f"unexpected compressed tensors: {sorted(unknown)}")
        if "weight_packed" not in entries and "weight" not in entries:
            raise ValueError("compressed entries need weight_packed or weight")
        if "weight_scale" not in entries:
            raise ValueError("compressed entries need weight_scale")
        for k, v in entries.items():
            self.register_buffer(k, v, persistent=True)
        self.register_buffer("bias", bias, persistent=bias is not None)
        self.args = args
        self.in_features, self.out_features = int(in_features), int(out_features)
        self.dtype = dtype
        self.cache = cache
        self._decoded: Optional[torch.Tensor] = None

    @classmethod
    def from_linear(cls, module: torch.nn.Linear, entries: Dict[str, torch.Tensor], args, cache: bool = False) -> "CompressedLinear":
        return cls(entries, args, module.in_features, module.out_features, None if module.bias is None else module.bias.detach(),
                   dtype=module.weight.dtype, cache=cache)

    def _buf(self, name: str) -> Optional[torch.Tensor]:
        return getattr(self, name, None)

    @torch.no_grad()
    def decompressed_weight(self) -> torch.Tensor:
        if self._decoded is not None:
            return self._decoded
        a = self.args
        shape = (self.out_features, self.in_features)
        packed, scale = self._buf("weight_packed"), self._buf("weight_scale")
        if a.type == "float" and a.num_bits == 4:
            w = ops.decompress_nvfp4(packed, scale, self._buf("weight_global_scale"), dtype=self.dtype)
        elif packed is not None:
            w = ops.decompress_int_packed(packed, scale, self._buf("weight_zero_point"), shape, a)
        else:
            w = ops.dequantize(self._buf("weight"), scale, self._buf("weight_zero_point"), args=a, dtype=self.dtype)
        if tuple(w.shape) != shape:
            raise RuntimeError(f"decoded weight has shape {tuple(w.shape)}, expected {shape}")
        if w.dtype != self.dtype:
            w = w.to(self.dtype)
        if self.cache:
            self._decoded = w
        return w

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.linear(x, self.decompressed_weight(), self.bias)

    def extra_repr(self) -> str:
        a = self.args
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, {a.type}{a.num_bits} {a.strategy}"


def _set_submodule(model: torch.nn.Module, name: str, new: torch.nn.Module) -> None:
    parent_name, _, leaf = name.rpartition(".")
    parent = model.get_submodule(parent_name) if parent_name else model
    setattr(parent, leaf, new)


def apply_compressed(model: torch.nn.Module, recipe: Union[str, dict, Recipe], state_dict: Dict[str, torch.Tensor], cache: bool = False) -> List[str]:
    """Swap every Linear the recipe quantized for a ``CompressedLinear`` over its entries of ``state_dict`` (the dict ``oneshot``
    returned, or tensors loaded from a compressed checkpoint).  The dense ``weight`` of those modules is released.  Returns the
    swapped module names."""
    from .oneshot import _as_recipe, _linears

    rec = _as_recipe(recipe)
    lin = _linears(model)
    modules = dict(lin)
    swapped: List[str] = []
    for spec in rec.modifiers:
        if not spec.config_groups:
            continue
        for name, g in resolve_targets(lin, spec).items():
            if g.weights is None or name in swapped:
                continue
            entries = {k: state_dict[f"{name}.{k}"] for k in _KEYS if f"{name}.{k}" in state_dict}
            if "weight_scale" not in entries:
                continue  # targeted by the recipe but absent from this state dict (e.g. another rank's shard)
            _set_submodule(model, name, CompressedLinear.from_linear(modules[name], entries, g.weights, cache=cache))
            swapped.append(name)
    return swapped
