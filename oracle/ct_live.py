"""Drive the LIVE ``compressed_tensors`` package (the reference's own arithmetic) -- TEST INFRASTRUCTURE ONLY.

``mratsim/quantizers`` has no arithmetic of its own: ``scripts/do_oneshot.py:179-197`` hands the model to
``llmcompressor.oneshot`` and ``save_pretrained(save_compressed=True)``, which run compressed-tensors (CT).
CT 0.15.0.1 is installed in this image, llmcompressor is not; this module therefore runs the real CT calls
(``calculate_qparams``, ``generate_gparam``, ``Compressor.compress``, ``fake_quantize``) and restates only the
llmcompressor glue around them (observer flatten + amin/amax, update_weight_zp_scale's copy into the
module Parameters) following SURVEY.md Appendix A.

Used by: tests/golden/make_golden.py (fixture generation), tests/test_oracle_vs_ct.py (live cross-check),
bench.py ``--impl reference`` / ``cpu_baseline`` (the reference CPU path timed on the host cores).
Always run with TORCHDYNAMO_DISABLE=1: FP4_E2M1_DATA.cast_to_fp4 is @torch.compile and CPU Inductor is
not usable here (CT:quantization/quant_args.py:53-54).
"""
from __future__ import annotations

import os

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import torch  # noqa: E402


def available() -> bool:
    try:
        import compressed_tensors  # noqa: F401

        return True
    except Exception:
        return False


def make_args(qtype: str, num_bits: int, symmetric: bool, strategy: str, group_size=None, block_structure=None):
    from compressed_tensors.quantization import QuantizationArgs

    kw = dict(num_bits=num_bits, type=qtype, symmetric=symmetric, strategy=strategy, observer="memoryless_minmax")
    if group_size is not None:
        kw["group_size"] = group_size
    if block_structure is not None:
        kw["block_structure"] = list(block_structure)
    if qtype == "float" and num_bits == 4:
        kw["scale_dtype"] = torch.float8_e4m3fn
        kw["zp_dtype"] = torch.float8_e4m3fn
    return QuantizationArgs(**kw)


def flatten_weight(w: torch.Tensor, args) -> torch.Tensor:
    """llmcompressor observers/helpers.py::flatten_for_calibration (weights), restated (Appendix A)."""
    s = args.strategy
    if s == "tensor":
        return w.reshape(1, 1, -1)
    if s == "channel":
        return w.unsqueeze(-2).unsqueeze(0)
    if s in ("group", "tensor_group"):
        return w.unflatten(-1, (-1, args.group_size)).unsqueeze(0)
    if s == "block":
        bh, bw = args.block_structure
        from compressed_tensors.quantization.utils import maybe_pad_tensor_for_block_quant

        w = maybe_pad_tensor_for_block_quant(w, (bh, bw))
        rb, cb = w.shape[0] // bh, w.shape[1] // bw
        return w.reshape(rb, bh, cb, bw).transpose(1, 2).flatten(-2).unsqueeze(0)
    raise ValueError(s)


def observe_minmax(w: torch.Tensor, args):
    """memoryless_minmax: amin/amax over dims (0,-1) of the flattened weight."""
    obs = flatten_weight(w, args)
    return torch.amin(obs, dim=(0, -1)), torch.amax(obs, dim=(0, -1))


def global_scale(w: torch.Tensor) -> torch.Tensor:
    """Observer.get_global_scale: reshape(1,1,-1) -> min/max -> generate_gparam."""
    from compressed_tensors.quantization.utils.helpers import generate_gparam

    obs = w.reshape(1, 1, -1)
    mn, mx = torch.amin(obs, dim=(0, -1)), torch.amax(obs, dim=(0, -1))
    return generate_gparam(mn.reshape(1), mx.reshape(1))


def weight_qparams(w: torch.Tensor, args, gs: torch.Tensor | None = None):
    """update_weight_zp_scale: observer -> calculate_qparams -> copy_ into the module Parameters
    (scale in the weight dtype, zp in args.zp_dtype; CT:quantization/lifecycle/initialize.py:231-254)."""
    from compressed_tensors.quantization.utils.helpers import calculate_qparams

    mn, mx = observe_minmax(w, args)
    scale, zp = calculate_qparams(mn, mx, args, global_scale=gs)
    scale_p = torch.empty(scale.shape, dtype=w.dtype, device=w.device).copy_(scale)
    zp_p = torch.zeros(zp.shape, dtype=args.zp_dtype, device=w.device).copy_(zp)
    if args.strategy == "channel":
        scale_p, zp_p = scale_p.reshape(-1, 1), zp_p.reshape(-1, 1)
    return scale_p, zp_p


def compress(w: torch.Tensor, fmt: str, args, gs: torch.Tensor | None = None):
    """Observer + qparams + ``BaseCompressor.compress`` for one Linear weight; returns the CT state dict."""
    from compressed_tensors.compressors.base import BaseCompressor
    from compressed_tensors.quantization import QuantizationScheme

    scheme = QuantizationScheme(targets=["Linear"], weights=args)
    sd = {"weight": w}
    if args.strategy == "tensor_group":
        if gs is None:
            gs = global_scale(w)
        sd["weight_global_scale"] = gs
    scale, zp = weight_qparams(w, args, gs)
    sd["weight_scale"] = scale
    sd["weight_zero_point"] = zp
    comp = BaseCompressor.get_value_from_registry(fmt)
    return comp.compress(sd, scheme)


def fake_quantize(w, scale, zp, args, gs=None):
    from compressed_tensors.quantization.lifecycle.forward import fake_quantize as fq

    return fq(w, scale, zp, args, global_scale=gs)


def mse_minmax(w: torch.Tensor, args, gs: torch.Tensor | None = None, maxshrink: float = 0.2, patience: int = 5, grid: float = 100.0,
               norm: float = 2.4):
    """llmcompressor observers/mse.py::_grid_search_mse restated on the LIVE compressed-tensors calls (calculate_qparams,
    fake_quantize with the strategy patched to TOKEN, as upstream does): only the loop and the reductions are re-typed."""
    import copy

    from compressed_tensors.quantization import QuantizationStrategy
    from compressed_tensors.quantization.utils.helpers import calculate_qparams

    observed = flatten_weight(w, args)
    min_val, max_val = torch.amin(observed, dim=(0, -1)), torch.amax(observed, dim=(0, -1))
    best_error = torch.full_like(min_val, torch.finfo(min_val.dtype).max)
    best_min, best_max = min_val.clone(), max_val.clone()
    token_args = copy.deepcopy(args)
    token_args.strategy = QuantizationStrategy.TOKEN
    no_improve = 0
    for i in range(int(maxshrink * grid)):
        p = 1 - i / grid
        smin, smax = p * min_val, p * max_val
        scales, zps = calculate_qparams(min_vals=smin, max_vals=smax, quantization_args=args, global_scale=gs)
        q = fake_quantize(observed, scales.unsqueeze(-1), zps.unsqueeze(-1), token_args, gs).to(observed.dtype)
        q -= observed
        q.abs_()
        q.pow_(norm)
        err = torch.sum(q, dim=(0, -1))
        better = err < best_error
        if torch.any(better):
            best_error[better] = err[better]
            best_min[better], best_max[better] = smin[better], smax[better]
            no_improve = 0
        else:
            no_improve += 1
            if no_improve >= patience:
                break
    return best_min, best_max


# ----------------------------------------------------------------------------- canonical format table
FORMATS = {
    # name: (compressor format, qtype, num_bits, symmetric, strategy, group, block)
    "int4_g128_asym": ("pack-quantized", "int", 4, False, "group", 128, None),
    "int4_g128_sym": ("pack-quantized", "int", 4, True, "group", 128, None),
    "int4_g32_sym": ("pack-quantized", "int", 4, True, "group", 32, None),
    "int4_g32_asym": ("pack-quantized", "int", 4, False, "group", 32, None),
    "int4_channel_sym": ("pack-quantized", "int", 4, True, "channel", None, None),
    "int4_channel_asym": ("pack-quantized", "int", 4, False, "channel", None, None),
    "int8_g128_sym": ("pack-quantized", "int", 8, True, "group", 128, None),
    "int8_channel_sym": ("pack-quantized", "int", 8, True, "channel", None, None),
    "fp8_channel": ("float-quantized", "float", 8, True, "channel", None, None),
    "fp8_g32": ("float-quantized", "float", 8, True, "group", 32, None),
    "fp8_g128": ("float-quantized", "float", 8, True, "group", 128, None),
    "fp8_block": ("float-quantized", "float", 8, True, "block", None, (128, 128)),
    "fp8_tensor": ("float-quantized", "float", 8, True, "tensor", None, None),
    "nvfp4": ("nvfp4-pack-quantized", "float", 4, True, "tensor_group", 16, None),
}


def format_args(name: str):
    fmt, qtype, nb, sym, strat, g, blk = FORMATS[name]
    return fmt, make_args(qtype, nb, sym, strat, g, blk)
