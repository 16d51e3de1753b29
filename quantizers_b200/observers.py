"""CUDA-backed calibration observers under the reference's registry names.

llmcompressor selects an observer from the recipe by name (``observer: memoryless_minmax`` in
/root/reference/configs/recipes/recipe_awq_w4a16.yaml:24; ``minmax`` in recipe_Minimax-M2.1-AWQ-MixedPrec.yaml:48;
``static_minmax`` is the NVFP4 input default, CT:quantization/quant_scheme.py:170-180) and calls

    observer = Observer.load_from_registry(name, base_name="weight", args=quantization_args, module=module)
    scale, zero_point = observer(observed)              # Observer.forward
    global_scale      = observer.get_global_scale(observed)

(LLMC observers/base.py, observers/min_max.py, observers/moving_base.py, restated in SURVEY.md Appendix A).  The classes
below keep that constructor / method surface; the reductions (``flatten_for_calibration`` + amin/amax) and
``calculate_qparams`` / ``generate_gparam`` run in the CUDA kernels behind ``ops.observe_minmax`` /
``ops.observe_global_scale`` / ``ops.calculate_qparams``.  Observer state (running or averaged min/max) stays on the GPU.
With token-sharded calibration ``sync()`` all-reduces static min/max state (MIN / MAX are order independent, so the
result is bit-identical for any world size); the EMA observer is order dependent and refuses to be sharded.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Type

import torch

from . import _lib as L
from . import ops

_REGISTRY: Dict[str, Type["Observer"]] = {}


def _strategy(args) -> str:
    return getattr(args.strategy, "value", args.strategy)


class Observer(torch.nn.Module):
    """Base class: ``forward(observed) -> (scale, zero_point)``; subclasses define how new min/max fold into the state."""

    def __init__(self, base_name: str = "weight", args=None, module: Optional[torch.nn.Module] = None, **observer_kwargs):
        super().__init__()
        self.base_name, self.args, self.module = base_name, args, module
        self.kwargs = {**(getattr(args, "observer_kwargs", None) or {}), **observer_kwargs}
        self.min_val: Optional[torch.Tensor] = None
        self.max_val: Optional[torch.Tensor] = None
        self.global_min: Optional[torch.Tensor] = None
        self.global_max: Optional[torch.Tensor] = None

    # ---- registry (same decorator / lookup shape as LLMC's RegistryMixin)
    @classmethod
    def register(cls, name: str):
        def deco(klass):
            _REGISTRY[name] = klass
            klass.registry_name = name
            return klass

        return deco

    @classmethod
    def load_from_registry(cls, name: str, **kwargs) -> "Observer":
        if name not in _REGISTRY:
            raise KeyError(f"Unable to find observer {name!r} in the registry; known: {sorted(_REGISTRY)}")
        return _REGISTRY[name](**kwargs)

    @classmethod
    def registered_names(cls):
        return sorted(_REGISTRY)

    # ---- statistics of one observation
    def _observe(self, observed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        L.require_cuda(observed)
        strat = _strategy(self.args)
        if self.base_name == "weight":
            return ops.observe_minmax(observed, self.args)
        # activations [..., H]: static per-tensor quantization (TENSOR) and the TENSOR_GROUP *global* statistics are whole-
        # tensor reductions; dynamic (TOKEN / per-group) activation qparams are never calibrated
        if strat in ("tensor", "tensor_group"):
            return self._tensor_minmax(observed)
        raise L.B200QError(f"activation observers support per-tensor statistics; strategy {strat!r} is dynamic at run time")

    @staticmethod
    def _tensor_minmax(observed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        x = observed.contiguous()
        state = torch.empty(2, dtype=torch.float32, device=x.device)
        L.check(L.lib().b200q_global_scale(L.ptr(x), 1, x.numel(), L.DTYPE_CODE[x.dtype], L.ptr(state), 0, None, L.stream_ptr(x.device)))
        return state[0:1].to(x.dtype), state[1:2].to(x.dtype)

    def _update(self, prev_min, prev_max, mn, mx):  # -> new (min, max)
        raise NotImplementedError

    def get_min_max(self, observed: torch.Tensor):
        mn, mx = self._observe(observed)
        self.min_val, self.max_val = self._update(self.min_val, self.max_val, mn, mx)
        return self.min_val, self.max_val

    def get_global_min_max(self, observed: torch.Tensor):
        mn, mx = self._tensor_minmax(observed)
        self.global_min, self.global_max = self._update(self.global_min, self.global_max, mn, mx)
        return self.global_min, self.global_max

    @torch.no_grad()
    def forward(self, observed: torch.Tensor):
        if observed.numel() == 0:  # unrouted expert: calibrate_activations skips empty inputs
            raise L.B200QError("cannot observe an empty tensor")
        gs = getattr(self.module, f"{self.base_name}_global_scale", None) if self.module is not None else None
        mn, mx = self.get_min_max(observed)
        return ops.calculate_qparams(mn, mx, self.args, global_scale=gs)

    @torch.no_grad()
    def get_global_scale(self, observed: torch.Tensor) -> torch.Tensor:
        mn, mx = self.get_global_min_max(observed)
        return ops.generate_gparam(mn, mx)

    def reset(self):
        self.min_val = self.max_val = self.global_min = self.global_max = None

    def sync(self, group=None):
        """Token-sharded calibration: fold the other ranks' statistics into this observer's state."""
        raise L.B200QError(f"{type(self).__name__} state cannot be combined across token shards")


@Observer.register("memoryless_minmax")
class MemorylessMinMaxObserver(Observer):
    """No state: every call stands alone (default for static weights, CT:quantization/quant_args.py:376-378)."""

    def _update(self, prev_min, prev_max, mn, mx):
        return mn, mx

    def sync(self, group=None):
        _allreduce_minmax(self, group)


@Observer.register("static_minmax")
class StaticMinMaxObserver(Observer):
    """Running element-wise min / max over all observations (NVFP4 ``input_global_scale``)."""

    def _update(self, prev_min, prev_max, mn, mx):
        if prev_min is None:
            return mn, mx
        return torch.minimum(prev_min, mn), torch.maximum(prev_max, mx)

    def sync(self, group=None):
        _allreduce_minmax(self, group)


@Observer.register("minmax")
class MovingAverageMinMaxObserver(Observer):
    """Exponential moving average ``past + c (cur - past)``, ``averaging_constant`` c = 0.01; the first observation (or
    c == 1) is taken as is (LLMC observers/moving_base.py)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.averaging_constant = float(self.kwargs.get("averaging_constant", 0.01))

    def _update(self, prev_min, prev_max, mn, mx):
        c = self.averaging_constant
        if prev_min is None or c == 1.0:
            return mn, mx
        return prev_min + c * (mn - prev_min), prev_max + c * (mx - prev_max)


class _MSEMixin:
    """Shrink-grid MSE range search (LLMC observers/mse.py ``_grid_search_mse``; maxshrink 0.2, patience 5, grid 100, norm 2.4,
    overridable through ``observer_kwargs``) in place of the plain amin / amax: ``ops.observe_mse_minmax`` evaluates all grid
    points of every quantization chunk in one pass over the weight.  Global (per-tensor) statistics stay plain min / max."""

    def _observe(self, observed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        L.require_cuda(observed)
        if self.base_name != "weight":
            raise NotImplementedError("mse observers are implemented for weights (GROUP / TENSOR_GROUP / CHANNEL)")
        gs = getattr(self.module, f"{self.base_name}_global_scale", None) if self.module is not None else None
        kw = self.kwargs
        return ops.observe_mse_minmax(observed, self.args, global_scale=gs, maxshrink=float(kw.get("maxshrink", 0.2)),
                                      patience=int(kw.get("patience", 5)), grid=int(float(kw.get("grid", 100.0))),
                                      norm=float(kw.get("norm", 2.4)))


@Observer.register("memoryless_mse")
class MemorylessMSEObserver(_MSEMixin, MemorylessMinMaxObserver):
    """MSE-selected range, no state."""

    def sync(self, group=None):
        raise L.B200QError("mse ranges are not min/max statistics: they cannot be combined across shards")


@Observer.register("mse")
class MovingAverageMSEObserver(_MSEMixin, MovingAverageMinMaxObserver):
    """MSE-selected range folded into an exponential moving average (first observation taken as is)."""


def _allreduce_minmax(obs: Observer, group=None):
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for lo, hi in (("min_val", "max_val"), ("global_min", "global_max")):
        a, b = getattr(obs, lo), getattr(obs, hi)
        if a is None:
            continue
        fa, fb = a.float(), b.float()  # NCCL has no bf16 MIN/MAX on every version; fp32 round trip is exact
        dist.all_reduce(fa, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(fb, op=dist.ReduceOp.MAX, group=group)
        setattr(obs, lo, fa.to(a.dtype))
        setattr(obs, hi, fb.to(b.dtype))
