"""Install the CUDA implementations at the reference's seams.

* compressor registry: ``BaseCompressor.get_value_from_registry(format)`` resolves to the B200 compressors;
* name-binding sites: ``quantize`` / ``dequantize`` / ``fake_quantize`` / ``calculate_qparams`` / ``generate_gparam`` /
  pack helpers are imported *by name* into several compressed-tensors modules, so each binding is replaced.

    with quantizers_b200.patch.patch():
        model.save_pretrained(out, save_compressed=True)      # /root/reference/scripts/do_oneshot.py:197

``patch()`` also works as a plain call (``handle = patch(); ...; handle.restore()``).
"""
from __future__ import annotations

import importlib
from contextlib import AbstractContextManager

from . import ops

# (module, attribute) -> replacement
_BINDINGS = [
    ("compressed_tensors.quantization.lifecycle.forward", "quantize", ops.quantize),
    ("compressed_tensors.quantization.lifecycle.forward", "dequantize", ops.dequantize),
    ("compressed_tensors.quantization.lifecycle.forward", "fake_quantize", ops.fake_quantize),
    ("compressed_tensors.quantization.lifecycle.compressed", "quantize", ops.quantize),
    ("compressed_tensors.compressors.pack_quantized.base", "quantize", ops.quantize),
    ("compressed_tensors.compressors.pack_quantized.base", "dequantize", ops.dequantize),
    ("compressed_tensors.compressors.pack_quantized.base", "pack_to_int32", ops.pack_to_int32),
    ("compressed_tensors.compressors.pack_quantized.base", "unpack_from_int32", ops.unpack_from_int32),
    ("compressed_tensors.compressors.nvfp4.base", "quantize", ops.quantize),
    ("compressed_tensors.compressors.nvfp4.base", "dequantize", ops.dequantize),
    ("compressed_tensors.compressors.nvfp4.base", "pack_fp4_to_uint8", ops.pack_fp4_to_uint8),
    ("compressed_tensors.compressors.nvfp4.base", "unpack_fp4_from_uint8", ops.unpack_fp4_from_uint8),
    ("compressed_tensors.compressors.naive_quantized.base", "quantize", ops.quantize),
    ("compressed_tensors.compressors.naive_quantized.base", "dequantize", ops.dequantize),
    ("compressed_tensors.quantization.utils.helpers", "calculate_qparams", ops.calculate_qparams),
    ("compressed_tensors.quantization.utils.helpers", "generate_gparam", ops.generate_gparam),
    ("compressed_tensors.quantization.utils", "calculate_qparams", ops.calculate_qparams),
    ("compressed_tensors.quantization.utils", "generate_gparam", ops.generate_gparam),
    # llmcompressor (absent in this image; patched when importable) binds these by name as well
    ("llmcompressor.observers.base", "calculate_qparams", ops.calculate_qparams),
    ("llmcompressor.observers.base", "generate_gparam", ops.generate_gparam),
    ("llmcompressor.modifiers.awq.base", "forward_quantize", None),  # resolved through forward.fake_quantize above
]


class patch(AbstractContextManager):
    def __init__(self, compressors: bool = True, functions: bool = True):
        self._saved = []
        self._saved_registry = {}
        if functions:
            for mod_name, attr, repl in _BINDINGS:
                if repl is None:
                    continue
                try:
                    mod = importlib.import_module(mod_name)
                except Exception:
                    continue
                if hasattr(mod, attr):
                    self._saved.append((mod, attr, getattr(mod, attr)))
                    setattr(mod, attr, repl)
        if compressors:
            from compressed_tensors.compressors.base import BaseCompressor
            from compressed_tensors.registry.registry import _REGISTRY

            from .compressors import REGISTRY_OVERRIDES

            reg = _REGISTRY[BaseCompressor]
            for name, klass in REGISTRY_OVERRIDES.items():
                if name in reg:
                    self._saved_registry[name] = reg[name]
                    reg[name] = klass

    def restore(self):
        for mod, attr, orig in reversed(self._saved):
            setattr(mod, attr, orig)
        self._saved = []
        if self._saved_registry:
            from compressed_tensors.compressors.base import BaseCompressor
            from compressed_tensors.registry.registry import _REGISTRY

            _REGISTRY[BaseCompressor].update(self._saved_registry)
            self._saved_registry = {}

    def __exit__(self, *exc):
        self.restore()
        return False
