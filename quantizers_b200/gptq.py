"""GPTQ Hessian accumulation on the tensor cores.

llmcompressor's ``GPTQModifier`` (the reference's GPTQ recipes, REF:scripts/old_scripts/main_glm4-gptq.py:108-126) keeps per Linear

    H    <- H * n / (n + t)                       # n = samples seen so far, t = rows of this batch
    n    <- n + t
    inp  <- sqrt(2 / n) * x.float().T             # [features, t]
    H    <- H + inp @ inp.T

(LLMC modifiers/gptq/gptq_quantize.py ``accumulate_hessian``).  ``X^T X`` is a dense contraction, so it runs as one tcgen05 GEMM
with an fp32 read-modify-write epilogue: ``H = (n / (n + t)) * H + (2 / (n + t)) * X^T X`` -- the same quantity with the scale
applied after the fp32 accumulation instead of before it (bf16 products are exact in fp32 either way).  Only the Hessian
statistic is provided here; the sequential GPTQ weight update is outside the hot path this package accelerates.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib as L


@torch.no_grad()
def accumulate_hessian(x: torch.Tensor, hessian: torch.Tensor, num_samples: int) -> Tuple[torch.Tensor, int]:
    """x ``[..., features]`` bf16 (inputs of the Linear for one calibration batch), hessian fp32 ``[features, features]`` updated in
    place, ``num_samples`` the running row count.  Returns (hessian, new num_samples)."""
    L.require_cuda(x, hessian)
    if x.dtype != torch.bfloat16 or hessian.dtype != torch.float32 or not hessian.is_contiguous():
        raise L.B200QError("accumulate_hessian takes bf16 activations and a contiguous fp32 Hessian")
    x2 = x.reshape(-1, x.shape[-1])
    t, k = x2.shape
    if hessian.shape != (k, k):
        raise L.B200QError(f"hessian must be [{k}, {k}], got {tuple(hessian.shape)}")
    if t == 0:
        return hessian, num_samples
    pad = (-t) % 8                                  # 16-byte rows for the tensor maps; zero rows add nothing to X^T X
    xt = torch.zeros((k, t + pad), dtype=x2.dtype, device=x2.device) if pad else torch.empty((k, t), dtype=x2.dtype, device=x2.device)
    xt[:, :t].copy_(x2.t())
    n = num_samples + t
    L.check(L.lib().b200q_gptq_hessian_accumulate(L.ptr(xt), k, t + pad, float(num_samples) / float(n), 2.0 / float(n), L.ptr(hessian),
                                                  L.stream_ptr(x2.device)))
    return hessian, n
