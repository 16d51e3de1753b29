#!/usr/bin/env python
"""bench.py -- headline benchmark of the quantization hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): Qwen3-4B (random-init, synthetic) mixed FP8 128x128 block (attention
q/k/v/o) + INT4 W4A16 group-128 asymmetric (MLP gate/up/down), RTN quantize+pack of all 36 decoder layers
= 252 matrices = 7.27 GB of bf16 weights per step.  One step = one pass of the fused
observer -> qparams -> quantize -> pack path over the whole model.

  value   : bf16 weight GB/s quantized+packed, weights already resident in HBM (device-timed, CUDA events)
  e2e     : same metric through the C-ABI host pipeline (b200q_pipeline_compress_host): pinned HOST weights in,
            packed HOST tensors out, host<->device copies inside the timed region
  roofline: dominant kernel (INT4 group kernel) algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  N > 1   : weak scaling -- every rank quantizes its own 36-layer shard (layers are independent units; no
            data-path collective), value = N * bytes / max-over-ranks time.  ``headline_strong`` is the same job with a FIXED 36
            layers partitioned over the ranks (scheduler.partition), replayed as one CUDA graph per rank.
  legs    : further driver-visible legs -- config 1 AWQ search (layers/s, tensor roofline), config 3 GLM-4.7-Flash FP8 block /
            per-channel + model_free_ptq file -> file, config 4 NVFP4 experts (strong), config 5 MiniMax expert mappings (strong);
            a compact {value, frac} summary of every leg and the sharded == unsharded ``parity`` flags are the LAST keys of the line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "bf16_weight_GBps_quantized_packed"
UNIT = "GB/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.marks = {}
        self.proc = None
        self.index = index
        self._t = None

    def start(self):
        # NVML in-process (nvidia_ml_py): a sample costs ~0.1 ms, so even a 30 ms timed region (20 steps) gets tens of samples;
        # `nvidia-smi -lms` needs ~0.5 s to print its first line and often missed the region altogether.  Same counters as the recipe's
        # nvidia-smi line (clocks.sm, clocks.max.sm, power.draw, clocks_event_reasons.*); nvidia-smi stays as the fallback.
        try:
            import pynvml as N

            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            self._stop = threading.Event()

            def rd():
                while not self._stop.is_set():
                    try:
                        sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                        r = int(get_reasons(h))
                        pw = N.nvmlDeviceGetPowerUsage(h) / 1000.0
                    except Exception:
                        break
                    f = [str(sm), str(mx), f"{pw:.1f}"] + ["Active" if r & bits[n] else "Not Active"
                                                          for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
                    self.samples.append((time.time(), ", ".join(f)))
                    time.sleep(0.001)

            self._t = threading.Thread(target=rd, daemon=True)
            self._t.start()
            return
        except Exception:
            self._stop = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("B200Q_BENCH_LMS", "50"),
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def rd():
            for line in self.proc.stdout:
                self.samples.append((time.time(), line.strip()))

        self._t = threading.Thread(target=rd, daemon=True)
        self._t.start()

    def mark(self, name):
        self.marks[name] = time.time()

    def stop(self):
        if getattr(self, "_stop", None) is not None:
            self._stop.set()
            if self._t is not None:
                self._t.join(timeout=1)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        t0, t1 = self.marks.get("t0", 0), self.marks.get("t1", float("inf"))
        rows = [s for s in self.samples if t0 <= s[0] <= t1] or [s for s in self.samples if s[0] >= self.marks.get("w0", 0)]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if getattr(self, "_stop", None) is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------- reference / CPU baseline
def cpu_layer(layer_idx: int):
    """One Qwen3-4B decoder layer (7 matrices, 201.9 MB bf16) on the CPU, SURVEY.md §8d seeds."""
    shapes = [("q_proj", 4096, 2560, "fp8_block"), ("k_proj", 1024, 2560, "fp8_block"), ("v_proj", 1024, 2560, "fp8_block"),
              ("o_proj", 2560, 4096, "fp8_block"), ("gate_proj", 9728, 2560, "int4_g128_asym"),
              ("up_proj", 9728, 2560, "int4_g128_asym"), ("down_proj", 2560, 9728, "int4_g128_asym")]
    out = []
    for mi, (name, r, c, fmt) in enumerate(shapes):
        g = torch.Generator().manual_seed(1234 + layer_idx * 1000 + mi)
        out.append((name, fmt, (torch.randn(r, c, generator=g) * 0.02).to(torch.bfloat16)))
    return out


def cpu_quantize_layer(layer, kind: str):
    """The reference CPU path for one layer: live compressed_tensors (kind 'reference') or the C oracle ('port')."""
    nbytes = 0
    if kind == "reference":
        from oracle import ct_live as L

        for _, fmt_name, w in layer:
            fmt, args = L.format_args(fmt_name)
            L.compress(w, fmt, args)
            nbytes += w.numel() * 2
    else:
        from oracle import oracle as O

        for _, fmt_name, w in layer:
            if fmt_name == "fp8_block":
                O.compress(w, "float-quantized", O.Geom(O.BLOCK, 0, 128, 128), 8, True)
            else:
                O.compress(w, "pack-quantized", O.Geom(O.GROUP, 128), 4, False)
            nbytes += w.numel() * 2
    return nbytes


def cpu_kind():
    from oracle import ct_live as L

    return "reference" if L.available() else "port"


def cpu_observer_layer(layer, kind: str):
    """Observer-only part of the metric on the CPU: min/max statistics -> calculate_qparams for every matrix of one layer."""
    nbytes = 0
    for _, fmt_name, w in layer:
        if kind == "reference":
            from oracle import ct_live as L

            _, a = L.format_args(fmt_name)
            L.weight_qparams(w, a)
        else:
            from oracle import oracle as O

            geom = O.Geom(O.BLOCK, 0, 128, 128) if fmt_name == "fp8_block" else O.Geom(O.GROUP, 128)
            mn, mx = O.minmax(w, geom)
            O.calculate_qparams(mn, mx, O.FP8 if fmt_name == "fp8_block" else O.INT, 8 if fmt_name == "fp8_block" else 4, fmt_name == "fp8_block")
        nbytes += w.numel() * 2
    return nbytes


def awq_cpu_layer(tokens: int, full_tokens: int):
    """The restated llmcompressor search on LIVE compressed-tensors arithmetic (oracle/llmc_live.py), whole Qwen3-4B decoder layer
    (q/k/v, gate/up, down mappings, n_grid 20, W4A16 g128 asym) on ``tokens`` calibration tokens, torch CPU GEMMs on all host
    threads.  Returns (seconds measured, seconds extrapolated to ``full_tokens``, timers): the weight fake-quantisation does not
    depend on the token count, statistics and parent forwards scale linearly with it."""
    from oracle import ct_live as L
    from oracle import llmc_live as V
    from quantizers_b200 import scheduler as S

    w, acts = S.synth_awq_layer(0, tokens, torch.device("cpu"))
    _, a = L.format_args("int4_g128_asym")
    timers = {}
    t0 = time.perf_counter()
    V.search_decoder_layer(w, acts, a, n_heads=32, n_kv=8, head_dim=128, seq_len=512, timers=timers)
    sec = time.perf_counter() - t0
    k = full_tokens / tokens
    est = timers.get("weights", 0.0) + k * (timers.get("forward", 0.0) + timers.get("stats", 0.0))
    est += max(sec - sum(timers.values()), 0.0)
    return sec, est, {k2: round(v, 2) for k2, v in timers.items()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (live compressed_tensors when importable, else the
    oracle port) on the host cores; each step = one decoder layer of the same workload (bounded sample).  The other two parts
    of the BASELINE metric ride along as ``legs``: observer-only (statistics -> qparams) and the AWQ search of one layer on a
    token-subsampled problem (restated loop on live compressed-tensors)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = cpu_kind()
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads() if kind == "reference" else 1
    layers = [cpu_layer(i % 2) for i in range(2)]
    for i in range(max(args.warmup, 1) if args.warmup else 0):
        cpu_quantize_layer(layers[i % 2], kind)
    t0 = time.perf_counter()
    nbytes = 0
    for i in range(args.steps):
        nbytes += cpu_quantize_layer(layers[i % 2], kind)
    dt = time.perf_counter() - t0
    v = nbytes / dt / 1e9
    sample = "one Qwen3-4B decoder layer (7 matrices, 201.9 MB bf16) per step"
    legs = {}
    t0 = time.perf_counter()
    nb = sum(cpu_observer_layer(layers[i % 2], kind) for i in range(4))
    legs["observer_only"] = {"value": nb / (time.perf_counter() - t0) / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "4 decoder layers: min/max statistics -> calculate_qparams only"}
    if kind == "reference" and args.cpu_awq_tokens > 0:
        try:
            sec, est, timers = awq_cpu_layer(args.cpu_awq_tokens, args.awq_tokens)
            legs["awq"] = {"value": 1.0 / est, "unit": "layers/s", "cores": cores, "kind": "reference (live compressed-tensors arithmetic, restated llmcompressor loop)",
                           "measured_s": sec, "tokens": args.cpu_awq_tokens, "timers_s": timers,
                           "sample": f"whole decoder layer (q/k/v, gate/up, down mappings, n_grid 20) on {args.cpu_awq_tokens} tokens took {sec:.1f} s; "
                                     f"forward + statistics scaled by {args.awq_tokens}/{args.cpu_awq_tokens}, weight fake-quantisation kept -> {est:.0f} s per layer"}
        except Exception as e:  # noqa: BLE001 -- informational leg
            legs["awq"] = {"unavailable": str(e)[:200]}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "qwen3-4b mixed FP8_BLOCK(attn)+INT4 g128 asym(MLP) RTN quantize+pack", "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "legs": legs,
    }), flush=True)


# ----------------------------------------------------------------------------- e2e through the host pipeline
def run_e2e(spec, arena, steps, warmup, device_index):
    """Pinned host weights -> b200q_pipeline_compress_host -> pinned host outputs, per matrix (252 jobs / step)."""
    import ctypes

    from quantizers_b200 import _lib as L
    from quantizers_b200 import ops
    from quantizers_b200.scheduler import PRESETS

    from quantizers_b200 import numa

    lib = L.lib()
    jobs = []
    h2d = d2h = 0
    max_bytes = 0
    # pinned staging buffers on the GPU's own NUMA node (first touch happens inside the allocation): with 8 ranks the far-socket
    # hop, not PCIe, is what limits the host pipeline otherwise
    with numa.near_device(device_index) as bound:
        jobs, h2d, d2h, max_bytes = _e2e_buffers(spec, arena, ops, PRESETS)
    handle = ctypes.c_void_p()
    L.check(lib.b200q_pipeline_create(ctypes.byref(handle), max_bytes, device_index))

    def step():
        for hw, rows, cols, sc, codes, scale, zp in jobs:
            L.check(lib.b200q_pipeline_compress_host(handle, L.ptr(hw), 1, rows, cols, ctypes.byref(sc), L.ptr(codes), L.ptr(scale),
                                                     L.ptr(zp), None))
        L.check(lib.b200q_pipeline_sync(handle))

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    submit = 0.0
    for _ in range(steps):
        ts = time.perf_counter()
        for hw, rows, cols, sc, codes, scale, zp in jobs:
            L.check(lib.b200q_pipeline_compress_host(handle, L.ptr(hw), 1, rows, cols, ctypes.byref(sc), L.ptr(codes), L.ptr(scale),
                                                     L.ptr(zp), None))
        submit += time.perf_counter() - ts
        L.check(lib.b200q_pipeline_sync(handle))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    lib.b200q_pipeline_destroy(handle)
    # ---- what bounds it (round-1 verdict, weak #7): the same pinned buffers through bare copies, no kernels, all ranks at once
    import torch.distributed as dist

    def bare(direction):
        dev_buf = arena[spec.matrices[-1].name]
        src = jobs[-1][0]
        n = 0
        torch.cuda.synchronize()
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
        tb = time.perf_counter()
        for hw, rows, cols, sc, codes, scale, zp in jobs:
            if direction == "h2d":
                dev_buf.view(-1)[: hw.numel()].copy_(hw.view(-1), non_blocking=True)
                n += hw.numel() * 2
            else:
                codes.view(-1).copy_(dev_buf.view(torch.uint8).view(-1)[: codes.numel() * codes.element_size()].view(codes.dtype), non_blocking=True)
                n += codes.numel() * codes.element_size()
        torch.cuda.synchronize()
        return n / (time.perf_counter() - tb) / 1e9

    breakdown = {"h2d_only_GBps": bare("h2d"), "d2h_only_GBps": bare("d2h"), "submit_ms_per_step": submit / steps * 1e3,
                 "step_ms": dt / steps * 1e3, "calls_per_step": len(jobs),
                 "note": "bare cudaMemcpyAsync of the same pinned buffers on every rank at once vs the pipelined step (H2D and D2H overlap in the "
                         "pipeline); submit = host time spent issuing the step's C calls"}
    return h2d * steps / dt / 1e9, h2d, d2h, dt, bound, breakdown


def _e2e_buffers(spec, arena, ops, PRESETS):
    jobs = []
    h2d = d2h = 0
    max_bytes = 0
    for m in spec.matrices:
        w = arena[m.name]
        a = PRESETS[m.preset]
        hw = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
        hw.copy_(w)
        n_mat, rows, cols = w.shape
        max_bytes = max(max_bytes, rows * cols * 2)
        sc = ops.scheme_from_args(a, w.dtype, True)
        if a.type == "int":
            codes = torch.empty((n_mat, rows, cols // 8), dtype=torch.int32, pin_memory=True)
            scale = torch.empty((n_mat, rows, cols // a.group_size), dtype=w.dtype, pin_memory=True)
            zp = None if a.symmetric else torch.empty((n_mat, -(-rows // 8), cols // a.group_size), dtype=torch.int32, pin_memory=True)
        else:
            codes = torch.empty((n_mat, rows, cols), dtype=torch.uint8, pin_memory=True)
            scale = torch.empty((n_mat, -(-rows // 128), -(-cols // 128)), dtype=w.dtype, pin_memory=True)
            zp = None
        for i in range(n_mat):
            jobs.append((hw[i], rows, cols, sc, codes[i], scale[i], None if zp is None else zp[i]))
            h2d += hw[i].numel() * 2
            d2h += codes[i].numel() * codes.element_size() + scale[i].numel() * 2 + (0 if zp is None else zp[i].numel() * 4)
    return jobs, h2d, d2h, max_bytes


# ----------------------------------------------------------------------------- AWQ search leg (BASELINE.json metric part 2)
def awq_cpu_sample(tokens: int = 512):
    """The restated reference search (oracle/llmc_restated.py on the pinned compressed-tensors arithmetic, torch CPU GEMMs) for
    the up_proj -> down_proj mapping of one Qwen3-4B layer on ``tokens`` calibration tokens; returns seconds."""
    from oracle import llmc_restated as R
    from oracle import oracle as O

    g = torch.Generator().manual_seed(4321)
    x = (torch.randn(tokens, 9728, generator=g) * (1 + 3 * torch.rand(9728, generator=g))).to(torch.bfloat16)
    w = (torch.randn(2560, 9728, generator=g) * 0.02).to(torch.bfloat16)
    t0 = time.perf_counter()
    R.compute_best_scale([x], [w], R.linear_parent, O.Geom(O.GROUP, 128), O.INT, 4, False)
    return time.perf_counter() - t0


def run_awq(args, dev, world, rank, peaks):
    """AWQ scale search (n_grid 20, duo_scaling, W4A16 g128 asym) of whole Qwen3-4B decoder layers, config 1 shapes:
    T = 64 x 512 tokens, mappings q/k/v (attention parent), gate/up (MLP parent), down (Linear parent).  Each rank searches
    its own layers (layer-sharded, weak scaling).  Device-timed with CUDA events; e2e adds the H2D copy of the layer's
    weights + calibration activations from pinned host memory and the D2H read of the best scales."""
    import torch.distributed as dist

    from quantizers_b200 import awq
    from quantizers_b200 import scheduler as S

    T = args.awq_tokens
    cfg = dict(n_heads=32, n_kv=8, head_dim=128, seq_len=512)
    qargs = S.PRESETS["W4A16_ASYM"]
    w, acts = S.synth_awq_layer(rank, T, dev)
    flops = awq.decoder_layer_flops(T, 2560, 9728, 32, 8, 128, 512)

    def layer(ww, aa):
        return awq.search_decoder_layer({k: v.clone() for k, v in ww.items()}, aa, qargs, **cfg)

    import gc

    gc.collect()                # the e2e leg's pinned staging buffers and device arenas are released here, not inside the timed region
    torch.cuda.empty_cache()
    for _ in range(3):  # warm-up: workspace growth, SDPA planning, kernel attributes
        layer(w, acts)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.awq_layers + 1)]
    e0.record()
    marks[0].record()
    for i in range(args.awq_layers):
        res = layer(w, acts)
        marks[i + 1].record()
    e1.record()
    torch.cuda.synchronize()
    log("awq per-layer ms:", [round(marks[i].elapsed_time(marks[i + 1]), 1) for i in range(args.awq_layers)])
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    # e2e: pinned host -> device every layer, best scales back on the host (compute_best_scale returns CPU tensors)
    hw = {k: v.cpu().pin_memory() for k, v in w.items()}
    ha = {k: v.cpu().pin_memory() for k, v in acts.items()}
    h2d = sum(v.numel() * v.element_size() for v in list(hw.values()) + list(ha.values()))
    d2h = sum(v[0].numel() * 4 for v in res.values())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.awq_layers, 2))
    for _ in range(n_e2e):
        dw = {k: v.to(dev, non_blocking=True) for k, v in hw.items()}
        da = {k: v.to(dev, non_blocking=True) for k, v in ha.items()}
        awq.search_decoder_layer(dw, da, qargs, **cfg)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / n_e2e], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    ms_layer = float(ms.item()) / args.awq_layers
    tf = flops / (ms_layer * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1500.0)))
    out = {"metric": "awq_search_layers_per_s", "value": world * 1e3 / ms_layer, "unit": "layers/s", "ms_per_layer": ms_layer,
           "layers_timed_per_gpu": args.awq_layers,
           "config": {"workload": "qwen3-4b decoder layer AWQ W4A16 g128 asym search, n_grid 20, duo_scaling, mappings qkv/gate_up/down",
                      "tokens": T, "seq_len": 512, "flops_per_layer": flops, "best_ratios": {k: v[1] for k, v in res.items()}},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)" if peaks else "fallback",
                        "frac_of_burst": tf / float(peaks.get("bf16_tflops", 1667.5)) if peaks else None,
                        "kernels": "awq_gemm_project_kernel / awq_gemm_loss_kernel (tcgen05, TMEM)"},
           "e2e": {"value": world * 1e3 / float(e2e_ms.item()), "unit": "layers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.cpu_awq_tokens > 0:
        torch.set_num_threads(os.cpu_count() or 1)
        try:
            sec, est, timers = awq_cpu_layer(args.cpu_awq_tokens, T)
            out["cpu_baseline"] = {"value": 1.0 / est, "unit": "layers/s", "cores": torch.get_num_threads(),
                                   "kind": "reference (live compressed-tensors arithmetic, restated llmcompressor loop)",
                                   "measured_s": sec, "timers_s": timers,
                                   "sample": f"whole decoder layer (q/k/v, gate/up, down mappings, n_grid 20) on {args.cpu_awq_tokens} tokens took {sec:.1f} s; "
                                             f"forward + statistics scaled by {T}/{args.cpu_awq_tokens}, weight fake-quantisation kept -> {est:.0f} s per layer"}
        except Exception as e:  # noqa: BLE001 -- informational leg
            out["cpu_baseline"] = {"unavailable": str(e)[:200]}
    return out


# ----------------------------------------------------------------------------- MoE legs (BASELINE.json configs[3], configs[4])
def run_moe_nvfp4(args, dev, world, rank, peaks):
    """configs[3]: Qwen3-30B-A3B NVFP4 RTN, the 128 experts of every layer sharded across the ranks (strong scaling: the job is
    ``--moe-layers`` layers x 128 experts whatever N is).  One step = fused |max| -> min(gate, up) global scale -> e4m3 group
    scales -> packed e2m1 codes for this rank's experts of all those layers (two launches: gate/up stack, down stack)."""
    import torch.distributed as dist

    from quantizers_b200 import scheduler as S

    experts = S.partition(128, world, rank)
    units = [l * 128 + e for l in range(args.moe_layers) for e in experts]
    spec = S.qwen3_30b_a3b(layers=1, experts=len(units))
    arena = S.build_arena(spec, units, dev)
    nbytes_total = args.moe_layers * 128 * spec.unit_bytes()
    outs = S.alloc_outputs(spec, arena)                  # caller-owned outputs: a step allocates nothing
    for _ in range(3):
        S.quantize_arena(spec, arena, out=outs)
    torch.cuda.synchronize()
    # one CUDA graph replay per step (two kernel launches + their sync-word memsets), like the other strong legs: at N = 8 a step is
    # 0.3 ms of kernels
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        S.quantize_arena(spec, arena, out=outs, concurrent=args.concurrent_classes)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = args.moe_steps
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    del graph
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    a = S.PRESETS["NVFP4"]
    alg = a.bytes_per_element() * (len(units) * spec.unit_elements())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = alg / (ms * 1e-3) / 1e9
    return {"metric": "bf16_weight_GBps_quantized_packed", "value": nbytes_total / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms,
            "scaling": "strong", "steps": steps,
            "config": {"workload": "qwen3-30b-a3b NVFP4 RTN quantize+pack, experts sharded across ranks", "layers": args.moe_layers,
                       "experts_per_layer": 128, "experts_per_gpu_per_layer": len(experts), "bytes_total": nbytes_total,
                       "l2": f"inputs ({len(units) * spec.unit_bytes() / 1e9:.2f} GB/step/GPU) larger than the 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "kernel": "nvfp4_fused2_kernel (|max| -> global scale -> compress, one launch per stack; one CUDA graph replay per step)",
                         "alg_bytes_per_element": a.bytes_per_element(), "note": "per-GPU figure of the slowest rank"}}


def run_moe_awq(args, dev, world, rank, peaks):
    """configs[4] kind (ii): MiniMax-M2.1 experts-only AWQ (INT4 g32 sym), the per-expert ``w3 -> w2`` mappings of one layer
    (T = 64 x 512 tokens, every expert sees all tokens) with the experts sharded across the ranks (strong scaling over
    ``--moe-awq-experts`` experts).  Each mapping is an independent n_grid-20 search on the fused tcgen05 loss GEMM."""
    import torch.distributed as dist

    from quantizers_b200 import awq
    from quantizers_b200 import scheduler as S

    T = args.awq_tokens
    qargs = S.PRESETS["INT4_G32_SYM"]
    experts = S.partition(args.moe_awq_experts, world, rank)
    w1, w3, w2, xs = S.synth_moe_awq_experts(0, experts, T, dev)
    del w1
    awq.search_expert_mappings(xs, w2.clone(), qargs, smooth_weight=w3.clone())  # full warm-up pass: kernels loaded, workspaces grown
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    copies = [(w2.clone(), w3.clone()) for _ in range(reps)]  # the search smooths the weights in place
    torch.cuda.synchronize()
    e0.record()
    res = []
    for a, b in copies:
        res = awq.search_expert_mappings(xs, a, qargs, smooth_weight=b)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    flops = awq.expert_mapping_flops(T, 1536, 3072) * args.moe_awq_experts
    tf = awq.expert_mapping_flops(T, 1536, 3072) * len(experts) / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1500.0)))
    out = {"metric": "awq_expert_mappings_per_s", "value": args.moe_awq_experts / (ms * 1e-3), "unit": "experts/s", "ms_total": ms,
           "scaling": "strong",
           "config": {"workload": "minimax-m2.1 experts-only AWQ INT4 g32 sym, per-expert w3->w2 mappings, n_grid 20, duo_scaling",
                      "experts": args.moe_awq_experts, "experts_per_gpu": len(experts), "tokens": T, "flops_total": flops,
                      "best_ratios_first4": [r[1] for r in res[:4]]},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                        "kernel": "awq_gemm_loss_kernel (tcgen05, TMEM)", "note": "per-GPU figure of the slowest rank"}}
    del xs, copies, w2, w3
    awq.workspace.release()
    torch.cuda.empty_cache()
    if args.moe_block_experts > 0:
        out["layer_mapping"] = run_moe_block(args, dev, world, rank, peak)
    return out


def run_moe_block(args, dev, world, rank, peak):
    """configs[4] kind (i): the layer-wide mapping post_attention_layernorm -> every expert's w1, w3 (one scale vector, parent =
    the routed sparse-MoE block, top-8 routing) on a layer of ``--moe-block-experts`` experts (default: all 256 of MiniMax-M2.1), strong
    scaling.  N > 1: EXPERT-parallel (``awq.search_moe_block_mapping_ep``): rank r holds the ascending expert range r of the layer and
    a token shard; the shards are all-gathered, every rank scales / fake-quantises / multiplies only its own experts, and the block
    output is accumulated along the ring 0 -> .. -> N-1 over NCCL send/recv so that the reference's per-expert bf16 ``index_add_``
    rounding sequence is kept bit for bit -- the real exchange step of this path.  ``--moe-block-token-sharded`` runs round 1's
    token-sharded variant instead (all experts on every rank; |x| sums and loss accumulators all-reduced)."""
    import torch.distributed as dist

    from quantizers_b200 import awq
    from quantizers_b200 import scheduler as S

    E, H, I = args.moe_block_experts, 3072, 1536
    K = min(8, E)
    T = args.awq_tokens
    qargs = S.PRESETS["INT4_G32_SYM"]
    ep = world > 1 and not args.moe_block_token_sharded and E >= world
    mine = S.partition(E, world, rank) if ep else range(E)
    units = list(mine)
    w1 = S.synth_stack(units, I, H, 0, dev)
    w3 = S.synth_stack(units, I, H, 1, dev)
    w2 = S.synth_stack(units, H, I, 2, dev)
    g = torch.Generator(device=dev).manual_seed(4321)
    x_all = (torch.randn(T, H, generator=g, device=dev) * (1 + 3 * torch.rand(H, generator=g, device=dev))).to(torch.bfloat16)
    router = torch.randn(E, H, generator=g, device=dev) * 0.02
    tok = S.partition(T, world, rank)
    x = x_all[tok.start:tok.stop].contiguous()
    p = torch.softmax(x.float() @ router.t(), dim=-1)
    topk_w, topk_idx = torch.topk(p, K, dim=-1)
    topk_w = topk_w / topk_w.sum(-1, keepdim=True)
    del x_all
    pg = dist.group.WORLD if world > 1 else None

    def search():
        if ep:
            return awq.search_moe_block_mapping_ep(x, w1, w3, w2, topk_idx, topk_w, qargs, expert_offset=mine.start, process_group=pg)
        return awq.search_moe_block_mapping(x, w1, w3, w2, topk_idx, topk_w, qargs, process_group=pg)

    search()   # warm-up pass
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s, ratio, losses = search()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    share = T // world if ep else len(tok)           # routed pairs a rank multiplies: all tokens x its experts == its share of the pairs
    tf = awq.moe_block_flops(share, K, H, I) / (ms * 1e-3) / 1e12
    awq.workspace.release()
    if world == 1:
        coll = None
    elif ep:
        coll = "all-gather of the token shards; ring send/recv of the running bf16 block output (21 passes); all-reduce of w_mean sums and [n_grid] losses"
    else:
        coll = "all-reduce(SUM) of |x| sums and [n_grid] loss accumulators"
    return {"metric": "awq_layer_mappings_per_s", "value": 1e3 / ms, "unit": "mappings/s", "ms": ms, "scaling": "strong",
            "config": {"workload": "minimax-m2.1 layer-wide mapping (post_attention_layernorm -> all experts' w1/w3), routed block parent",
                       "experts": E, "top_k": K, "tokens": T, "sharding": "experts (ring-ordered combine)" if ep else ("tokens" if world > 1 else None),
                       "experts_per_gpu": len(units), "tokens_per_gpu": T if ep else len(tok), "best_ratio": ratio, "collective": coll},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                         "kernel": "awq_gemm_project_kernel, grouped mode (tcgen05, TMEM): one launch per stage over all experts + bf16 combine kernel",
                         "note": "routed pairs sorted expert-major, rows padded to 128-row tiles; per-GPU figure of the slowest rank"}}

# ----------------------------------------------------------------------------- strong-scaling headline, GLM leg, parity
def run_headline_strong(args, dev, world, rank, peaks):
    """The headline job with a FIXED size: the matrices of the 36 decoder layers of Qwen3-4B partitioned over the ranks at (class,
    layer) granularity, balanced by elements (scheduler.partition_balanced; whole layers would be 5 / 4 per GPU at N = 8, i.e. at
    best 7.2x; no data-path collective).  One step = this rank's layers; the 7 launches of a step are captured
    once in a CUDA graph and replayed (at N = 8 a step is ~0.2 ms of kernels, launch-bound otherwise)."""
    import torch.distributed as dist

    from quantizers_b200 import scheduler as S

    spec = S.qwen3_4b(layers=1)
    mine = S.partition_balanced(spec, 36, world)[rank]        # {class: [layers]}: balanced by elements (whole layers: 5 / 4 at N = 8)
    arena = S.build_arena_classes(spec, mine, dev)
    outs = S.alloc_outputs(spec, arena)
    for _ in range(3):
        S.quantize_arena(spec, arena, out=outs)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        S.quantize_arena(spec, arena, out=outs, concurrent=args.concurrent_classes)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(args.steps, 20)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    total = 36 * S.qwen3_4b(layers=1).unit_bytes()
    alg = sum(S.PRESETS[m.preset].bytes_per_element() * m.rows * m.cols * m.per_unit * len(mine.get(m.name, ())) for m in spec.matrices)
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return {"metric": METRIC, "value": total / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "scaling": "strong", "steps": steps,
            "config": {"workload": "qwen3-4b mixed FP8_BLOCK + INT4 g128 asym, the 36 layers' matrices partitioned over the ranks by (class, layer), balanced by elements",
                       "matrices_per_gpu": {k: len(v) for k, v in mine.items()},
                       "launch": "one CUDA graph replay per step" + (", classes on concurrent streams" if args.concurrent_classes else "")},
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "note": "whole step (all 5 classes) of the slowest rank"}}


def run_glm(args, dev, world, rank, peaks):
    """configs[2]: GLM-4.7-Flash-shaped FP8 quantization (REF:scripts/quant_GLM-4.7-Flash-FP8.py:11-24).  (a) the fused kernels on the
    synthetic expert / dense shapes incl. the ragged [2624, 2048] rows (FP8_BLOCK 128x128, and FP8 per-channel = FP8_DYNAMIC weights),
    the ``--glm-units`` units sharded over the ranks (strong); (b) ``model_free_ptq`` file -> file on a page-cached synthetic
    safetensors checkpoint, shards split over the ranks: read -> pinned staging -> H2D -> kernel -> D2H -> pwrite."""
    import shutil
    import tempfile

    import torch.distributed as dist

    from quantizers_b200 import scheduler as S

    units = S.partition(args.glm_units, world, rank)
    peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {}
    for preset in ("FP8_BLOCK", "FP8_CHANNEL"):
        spec = S.glm47_flash(preset, units=len(units))
        arena = S.build_arena(spec, list(units), dev)
        outs = S.alloc_outputs(spec, arena)
        for _ in range(3):
            S.quantize_arena(spec, arena, out=outs)
        torch.cuda.synchronize()
        # one CUDA graph replay per step, like the strong headline leg: at N = 8 a step is ~0.1 ms of kernels in 4 launches
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            S.quantize_arena(spec, arena, out=outs, concurrent=args.concurrent_classes)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.moe_steps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        del graph
        ms = torch.tensor([e0.elapsed_time(e1) / args.moe_steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        total = args.glm_units * spec.unit_bytes()
        alg = S.PRESETS[preset].bytes_per_element() * len(units) * spec.unit_elements()
        out[preset.lower()] = {"metric": METRIC, "value": total / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "scaling": "strong",
                               "config": {"workload": f"glm-4.7-flash shapes {preset}: gate/up [1536,2048] x2, down [2048,1536], dense [10240,2048], "
                                                      "ragged [2624,2048] per unit", "units": args.glm_units, "units_per_gpu": len(units),
                                          "launch": "one CUDA graph replay per step"},
                               "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                            "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None}}
        del arena, outs
        torch.cuda.empty_cache()
    # ---- (b) model_free_ptq, file -> file
    if args.glm_file_gb > 0:
        from safetensors.torch import save_file

        from quantizers_b200.model_free import model_free_ptq

        root = None
        if rank == 0:
            root = tempfile.mkdtemp(prefix="b200q_glm_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        if world > 1:
            box = [root]
            dist.broadcast_object_list(box, src=0)
            root = box[0]
        src, dst = os.path.join(root, "src"), os.path.join(root, "dst")
        n_shards = max(8, world)
        per_layer = 2 * (1536 * 2048 * 2 + 2048 * 1536) + 10240 * 2048 * 2 + 2624 * 2048 * 2
        layers_per_shard = max(1, int(args.glm_file_gb * 1e9 / n_shards / per_layer))
        total = 0
        if rank == 0:
            os.makedirs(src)
            g = torch.Generator().manual_seed(0)
            for sh in range(n_shards):
                t = {}
                for i in range(layers_per_shard):
                    l = sh * layers_per_shard + i
                    t[f"model.layers.{l}.mlp.experts.0.gate_proj.weight"] = (torch.randn(1536, 2048, generator=g) * 0.02).to(torch.bfloat16)
                    t[f"model.layers.{l}.mlp.experts.0.up_proj.weight"] = (torch.randn(1536, 2048, generator=g) * 0.02).to(torch.bfloat16)
                    t[f"model.layers.{l}.mlp.experts.0.down_proj.weight"] = (torch.randn(2048, 1536, generator=g) * 0.02).to(torch.bfloat16)
                    t[f"model.layers.{l}.mlp.shared_experts.up_proj.weight"] = (torch.randn(10240, 2048, generator=g) * 0.02).to(torch.bfloat16)
                    t[f"model.layers.{l}.self_attn.kv_b_proj.weight"] = (torch.randn(2624, 2048, generator=g) * 0.02).to(torch.bfloat16)
                    t[f"model.layers.{l}.self_attn.q_a_proj.weight"] = (torch.randn(768, 2048, generator=g) * 0.02).to(torch.bfloat16)   # ignored
                    t[f"model.layers.{l}.input_layernorm.weight"] = torch.ones(2048, dtype=torch.bfloat16)
                save_file(t, os.path.join(src, f"model-{sh + 1:05d}-of-{n_shards:05d}.safetensors"), metadata={"format": "pt"})
            with open(os.path.join(src, "config.json"), "w") as f:
                json.dump({"model_type": "glm4_moe_lite", "hidden_size": 2048}, f)
        if world > 1:
            dist.barrier()
        total = sum(os.path.getsize(os.path.join(src, f)) for f in os.listdir(src) if f.endswith(".safetensors"))
        ignore = ["re:.*q_a_proj$", "re:.*kv_a_proj_with_mqa$", "re:.*mlp.gate$", "lm_head"]   # REF:scripts/quant_GLM-4.7-Flash-FP8.py:15-21
        kw = dict(scheme="FP8_BLOCK", ignore=ignore, max_workers=args.glm_workers, device=f"cuda:{dev.index}", rank=rank, world_size=world)
        model_free_ptq(src, dst + "_warm", **kw)      # warm-up: pinned arenas, pipeline handles, page cache
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        st = model_free_ptq(src, dst, **kw)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        out["model_free_ptq"] = {"metric": "bf16_checkpoint_GBps_file_to_file", "value": total / dt / 1e9, "unit": UNIT, "seconds": dt,
                                 "scaling": "strong",
                                 "config": {"workload": "model_free_ptq(scheme=FP8_BLOCK, ignore=[q_a_proj, kv_a_proj_with_mqa, mlp.gate, lm_head]) on a synthetic "
                                                        "GLM-4.7-Flash-shaped safetensors checkpoint in the page cache", "checkpoint_bytes": total,
                                            "shards": n_shards, "max_workers": args.glm_workers, "tensors_quantized_rank0": st["tensors_quantized"],
                                            "timing": "wall clock around the call (file read, H2D, kernel, D2H, pwrite), max over ranks"}}
        if world > 1:
            dist.barrier()
        if rank == 0:
            shutil.rmtree(root, ignore_errors=True)
    return out


def _checksum(t: torch.Tensor) -> torch.Tensor:
    """Order-independent exact checksum of a tensor's bytes (int64 sum of its 32-bit words / bytes)."""
    b = t.contiguous().view(torch.uint8).reshape(-1)
    n4 = b.numel() // 4 * 4
    return b[:n4].view(torch.int32).sum(dtype=torch.int64) + b[n4:].sum(dtype=torch.int64)


def run_parity(args, dev, world, rank):
    """Driver-visible correctness of the sharding (round-1 verdict, weak #9): the packed bytes a rank produces for ITS layers /
    experts equal what one GPU produces for the whole job, and the token-sharded AWQ search (NCCL all-reduce of |x| sums and loss
    accumulators) returns the unsharded argmin.  Checksums are all-reduced; rank 0 holds the unsharded reference."""
    import torch.distributed as dist

    from quantizers_b200 import awq
    from quantizers_b200 import ops
    from quantizers_b200 import scheduler as S

    res = {}
    # ---- RTN, 8 Qwen3-4B layers partitioned by layer
    n_layers = 8
    spec1 = S.qwen3_4b(layers=1)
    keys = ("weight_packed", "weight", "weight_scale", "weight_zero_point")
    sums = torch.zeros(n_layers, len(spec1.matrices), dtype=torch.int64, device=dev)
    mine = S.partition(n_layers, world, rank)
    if len(mine):
        spec = S.qwen3_4b(layers=len(mine))
        o = S.quantize_arena(spec, S.build_arena(spec, list(mine), dev))
        for mi, m in enumerate(spec.matrices):
            for li, layer in enumerate(mine):
                sl = slice(li * m.per_unit, (li + 1) * m.per_unit)
                sums[layer, mi] = sum(_checksum(o[m.name][k][sl]) for k in keys if k in o[m.name])
        del o
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if rank == 0:
        spec = S.qwen3_4b(layers=n_layers)
        o = S.quantize_arena(spec, S.build_arena(spec, list(range(n_layers)), dev))
        ref = torch.zeros_like(sums)
        for mi, m in enumerate(spec.matrices):
            for layer in range(n_layers):
                sl = slice(layer * m.per_unit, (layer + 1) * m.per_unit)
                ref[layer, mi] = sum(_checksum(o[m.name][k][sl]) for k in keys if k in o[m.name])
        res["rtn_layer_sharded_eq_unsharded"] = bool(torch.equal(ref, sums))
        del o
    # ---- the same 8 layers partitioned at (class, layer) granularity (the strong-scaling headline's partition)
    sums = torch.zeros(n_layers, len(spec1.matrices), dtype=torch.int64, device=dev)
    mine_c = S.partition_balanced(spec1, n_layers, world)[rank]
    if mine_c:
        o = S.quantize_arena(spec1, S.build_arena_classes(spec1, mine_c, dev))
        for mi, m in enumerate(spec1.matrices):
            for li, layer in enumerate(mine_c.get(m.name, ())):
                sl = slice(li * m.per_unit, (li + 1) * m.per_unit)
                sums[layer, mi] = sum(_checksum(o[m.name][k][sl]) for k in keys if k in o[m.name])
        del o
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if rank == 0:
        res["rtn_class_balanced_eq_unsharded"] = bool(torch.equal(ref, sums))
    # ---- NVFP4, 16 experts of one layer partitioned by expert (gate/up siblings stay together)
    n_exp = 16
    sums = torch.zeros(n_exp, 2, dtype=torch.int64, device=dev)
    mine = S.partition(n_exp, world, rank)
    keys4 = ("weight_packed", "weight_scale", "weight_global_scale")
    if len(mine):
        spec = S.qwen3_30b_a3b(layers=1, experts=len(mine))
        o = S.quantize_arena(spec, S.build_arena(spec, list(mine), dev))
        for mi, m in enumerate(spec.matrices):
            for ei, e in enumerate(mine):
                sl = slice(ei * m.per_unit, (ei + 1) * m.per_unit)
                sums[e, mi] = sum(_checksum(o[m.name][k][sl]) for k in keys4)
        del o
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if rank == 0:
        spec = S.qwen3_30b_a3b(layers=1, experts=n_exp)
        o = S.quantize_arena(spec, S.build_arena(spec, list(range(n_exp)), dev))
        ref = torch.zeros_like(sums)
        for mi, m in enumerate(spec.matrices):
            for e in range(n_exp):
                sl = slice(e * m.per_unit, (e + 1) * m.per_unit)
                ref[e, mi] = sum(_checksum(o[m.name][k][sl]) for k in keys4)
        res["nvfp4_expert_sharded_eq_unsharded"] = bool(torch.equal(ref, sums))
        del o
    # ---- AWQ, one single-Linear mapping, tokens sharded over the ranks
    T, K, N = 8192, 2560, 1024
    g = torch.Generator(device=dev).manual_seed(4242)
    x = (torch.randn(T, K, generator=g, device=dev) * (1 + 3 * torch.rand(K, generator=g, device=dev))).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
    tok = S.partition(T, world, rank)
    qa = S.PRESETS["W4A16_ASYM"]
    pg = dist.group.WORLD if world > 1 else None
    _, r_sh, l_sh = awq.compute_best_scale(x[tok.start:tok.stop].contiguous(), [w], awq.linear_parent, qa, process_group=pg)
    if rank == 0:
        _, r_un, l_un = awq.compute_best_scale(x, [w], awq.linear_parent, qa)
        res["awq_token_sharded_same_argmin"] = bool(r_sh == r_un)
        res["awq_token_sharded_max_rel_loss_diff"] = max(abs(a - b) / b for a, b in zip(l_sh, l_un))
    # ---- layer-wide MoE mapping, experts partitioned over the ranks, ring-ordered combine (16 experts, top-4, 4096 tokens)
    if world > 1:
        Em, Hm, Im, Km, Tm = 16, 512, 384, 4, 4096
        g = torch.Generator(device=dev).manual_seed(777)
        xm = (torch.randn(Tm, Hm, generator=g, device=dev) * (1 + 3 * torch.rand(Hm, generator=g, device=dev))).to(torch.bfloat16)
        m1 = (torch.randn(Em, Im, Hm, generator=g, device=dev) * 0.05).to(torch.bfloat16)
        m3 = (torch.randn(Em, Im, Hm, generator=g, device=dev) * 0.05).to(torch.bfloat16)
        m2 = (torch.randn(Em, Hm, Im, generator=g, device=dev) * 0.05).to(torch.bfloat16)
        pm = torch.softmax(torch.randn(Tm, Em, generator=g, device=dev), dim=-1)
        tw, ti = torch.topk(pm, Km, dim=-1)
        tw = tw / tw.sum(-1, keepdim=True)
        qm = S.PRESETS["INT4_G32_SYM"]
        ex, tk = S.partition(Em, world, rank), S.partition(Tm, world, rank)
        _, r_ep, l_ep = awq.search_moe_block_mapping_ep(xm[tk.start:tk.stop].contiguous(), m1[ex.start:ex.stop].contiguous(), m3[ex.start:ex.stop].contiguous(),
                                                        m2[ex.start:ex.stop].contiguous(), ti[tk.start:tk.stop].contiguous(), tw[tk.start:tk.stop].contiguous(),
                                                        qm, expert_offset=ex.start, process_group=pg)
        if rank == 0:
            _, r_1, l_1 = awq.search_moe_block_mapping(xm, m1, m3, m2, ti, tw, qm)
            res["moe_mapping_expert_parallel_same_argmin"] = bool(r_ep == r_1)
            res["moe_mapping_expert_parallel_max_rel_loss_diff"] = max(abs(a - b) / b for a, b in zip(l_ep, l_1))
    awq.workspace.release()
    torch.cuda.empty_cache()
    if rank == 0:
        if world > 1:
            res["ok_moe_mapping"] = bool(res["moe_mapping_expert_parallel_same_argmin"] and res["moe_mapping_expert_parallel_max_rel_loss_diff"] < 1e-3)
        res["ok"] = bool(res.get("ok_moe_mapping", True) and res["rtn_class_balanced_eq_unsharded"] and res["rtn_layer_sharded_eq_unsharded"] and res["nvfp4_expert_sharded_eq_unsharded"] and res["awq_token_sharded_same_argmin"]
                         and res["awq_token_sharded_max_rel_loss_diff"] < 1e-3)
        res["world_size"] = world
    return res


# ----------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=36, help="decoder layers per rank (36 = full Qwen3-4B)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-layers", type=int, default=36, help="decoder layers the CPU baseline quantizes (36 = the whole workload once)")
    ap.add_argument("--awq-layers", type=int, default=3, help="decoder layers of the AWQ search leg per rank (0 disables it)")
    ap.add_argument("--awq-tokens", type=int, default=64 * 512, help="calibration tokens per layer (64 samples x 512)")
    ap.add_argument("--moe-layers", type=int, default=8, help="layers of the Qwen3-30B-A3B NVFP4 expert-sharded leg (0 disables it)")
    ap.add_argument("--moe-steps", type=int, default=20)
    ap.add_argument("--moe-awq-experts", type=int, default=256, help="experts of the MiniMax-M2.1 per-expert AWQ leg (256 = one full layer; 0 disables it)")
    ap.add_argument("--concurrent-classes", type=int, default=1, help="strong legs: launch the independent matrix classes of a step on side streams inside the captured graph")
    ap.add_argument("--moe-block-token-sharded", action="store_true", help="N > 1: round 1's token-sharded layer-wide mapping instead of the expert-parallel ring")
    ap.add_argument("--moe-block-experts", type=int, default=256, help="experts of the layer-wide MoE mapping leg (256 = the full MiniMax-M2.1 layer; 0 disables it)")
    ap.add_argument("--cpu-awq-tokens", type=int, default=1024, help="calibration tokens of the CPU AWQ baseline (whole layer; 0 disables it); "
                                                                      "--impl reference uses max(this, 2048)")
    ap.add_argument("--glm-units", type=int, default=48, help="units of the GLM-4.7-Flash FP8 leg (0 disables it)")
    ap.add_argument("--glm-file-gb", type=float, default=2.0, help="size of the synthetic checkpoint of the model_free_ptq leg (0 disables it)")
    ap.add_argument("--glm-workers", type=int, default=4)
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling headline variant")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded == unsharded parity checks")
    args = ap.parse_args()
    if args.impl == "reference":
        args.cpu_awq_tokens = max(args.cpu_awq_tokens, 2048) if args.cpu_awq_tokens > 0 else 0
        run_reference(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the quantization hot path has no CPU fallback")
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)

    from quantizers_b200 import scheduler as S

    spec = S.qwen3_4b(layers=args.layers)
    units = list(range(rank * spec.units, (rank + 1) * spec.units))  # weak scaling: a distinct 36-layer shard per rank
    arena = S.build_arena(spec, units, dev)
    step_bytes = spec.units * spec.unit_bytes()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sampler.mark("w0")
    # caller-owned output + workspace buffers, allocated once: warm-up and timed steps are the same code writing the same buffers,
    # a step allocates nothing and never synchronises with the host (round 1 timed `out = quantize_arena(...)`: the second timed
    # step had to cudaMalloc a second 2.3 GB output set inside the gate_up launch's event pair -- 6..34 ms, profiles/r2_alloc_diag.md)
    outs = S.alloc_outputs(spec, arena)
    for _ in range(warmup):
        out = S.quantize_arena(spec, arena, [], out=outs)
    barrier()
    torch.cuda.synchronize()
    timings = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark("t0")
    e0.record()
    for _ in range(args.steps):
        out = S.quantize_arena(spec, arena, timings, out=outs)
    e1.record()
    torch.cuda.synchronize()
    sampler.mark("t1")
    barrier()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    sampler.stop()
    value = world * step_bytes * args.steps / (ms * 1e-3) / 1e9

    # ---- per-kernel-class times (CUDA events on the launching stream, inside the timed region)
    diag = os.environ.get("B200Q_BENCH_DIAG")
    if diag and rank == 0:
        nm = len(spec.matrices)
        rows = [[round(t[3].elapsed_time(t[4]), 4) for t in timings[i:i + nm]] for i in range(0, len(timings), nm)]
        gaps = [round(timings[i][3].elapsed_time(timings[i + nm][3]), 4) for i in range(0, len(timings) - nm, nm)]
        json.dump({"names": [m.name for m in spec.matrices], "per_step_ms": rows, "step_start_to_next_start_ms": gaps,
                   "total_ms": ms, "clock_samples": [(round(t - sampler.marks["t0"], 4), l) for t, l in sampler.samples],
                   "t1": sampler.marks["t1"] - sampler.marks["t0"]}, open(diag, "w"))
    per = {}
    for name, preset, elems, a, b in timings:
        d = per.setdefault(preset, {"ms": 0.0, "elems": 0, "n": 0})
        d["ms"] += a.elapsed_time(b)
        d["elems"] += elems
        d["n"] += 1
    by_name = {}
    for name, preset, elems, a, b in timings:
        d = by_name.setdefault(name, [0.0, 0, preset])
        d[0] += a.elapsed_time(b)
        d[1] += elems
    log("per-launch GB/s (algorithmic):", {k: round(S.PRESETS[v[2]].bytes_per_element() * v[1] / (v[0] * 1e-3) / 1e9) for k, v in by_name.items()})
    dom = max(per, key=lambda k: per[k]["ms"])
    dargs = S.PRESETS[dom]
    alg_bytes = dargs.bytes_per_element() * per[dom]["elems"]
    achieved = alg_bytes / (per[dom]["ms"] * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": ("group_tma_kernel<INT4, asym, g128> (TMA-staged fused observe+qparams+quantize+pack)" if dom == "W4A16_ASYM"
                           else f"fused compress kernel of {dom}"),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                "alg_bytes_per_element": dargs.bytes_per_element(), "avg_launch_ms": per[dom]["ms"] / per[dom]["n"],
                "per_class": {k: {"GBps_bf16_in": v["elems"] * 2 / (v["ms"] * 1e-3) / 1e9,
                                  "GBps_algorithmic": S.PRESETS[k].bytes_per_element() * v["elems"] / (v["ms"] * 1e-3) / 1e9,
                                  "share_of_step": v["ms"] / sum(x["ms"] for x in per.values())} for k, v in per.items()}}
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_file):
        try:  # ncu --set full: (dram__bytes_read.sum + dram__bytes_write.sum) / elements of the captured launch, scaled to this launch
            t = json.load(open(traffic_file)).get(dom)
            if t:
                roofline["traffic"] = t["dram_bytes_per_element"] * per[dom]["elems"] / per[dom]["n"]
                roofline["traffic_source"] = t["source"]
        except Exception:
            pass
    roofline["alg_bytes_per_launch"] = alg_bytes / per[dom]["n"]
    del out, outs

    # ---- e2e through the host pipeline (same metric, host buffers, copies in the timed region)
    e2e_v, h2d, d2h, _, numa_bound, e2e_breakdown = run_e2e(spec, arena, args.e2e_steps, 1, local)
    ev = torch.tensor([e2e_v], device=dev)
    if world > 1:
        dist.all_reduce(ev, op=dist.ReduceOp.MIN)  # slowest rank bounds the job
        e2e_v = float(ev.item()) * world
    barrier()

    del arena
    torch.cuda.empty_cache()
    strong = None if args.no_strong else run_headline_strong(args, dev, world, rank, peaks)
    barrier()
    parity = None if args.no_parity else run_parity(args, dev, world, rank)
    barrier()
    torch.cuda.empty_cache()
    glm = run_glm(args, dev, world, rank, peaks) if args.glm_units > 0 else None
    barrier()
    torch.cuda.empty_cache()
    awq_line = run_awq(args, dev, world, rank, peaks) if args.awq_layers > 0 else None
    barrier()
    from quantizers_b200 import awq as _awq

    _awq.workspace.release()
    torch.cuda.empty_cache()
    moe_nvfp4 = run_moe_nvfp4(args, dev, world, rank, peaks) if args.moe_layers > 0 else None
    barrier()
    torch.cuda.empty_cache()
    moe_awq = run_moe_awq(args, dev, world, rank, peaks) if args.moe_awq_experts > 0 else None
    barrier()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "qwen3-4b mixed FP8_BLOCK(attn)+INT4 g128 asym(MLP) RTN quantize+pack",
                   "layers_per_gpu": spec.units, "matrices_per_step": 7 * spec.units, "bytes_per_step_per_gpu": step_bytes,
                   "l2": "inputs (7.27 GB/step) far larger than the 126 MB L2; no flush needed", "parallelism": f"layer-sharded x{world}"},
        "roofline": roofline,
        "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                "host_buffers": "pinned, on the GPU's NUMA node" if numa_bound else "pinned", "breakdown": e2e_breakdown},
        "gpu_launches": S.launches_per_step(spec) * args.steps,
        "clocks": sampler.summary() if rank == 0 else None,
    }
    if strong is not None:
        line["headline_strong"] = strong
    if glm is not None:
        line["glm_fp8"] = glm
    if awq_line is not None:
        line["awq"] = awq_line
    if moe_nvfp4 is not None:
        line["moe_nvfp4"] = moe_nvfp4
    if moe_awq is not None:
        line["moe_awq"] = moe_awq
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            kind = cpu_kind()
            torch.set_num_threads(os.cpu_count() or 1)
            n_layers = max(1, args.cpu_layers)
            nb, dt = 0, 0.0
            for li in range(n_layers):          # generation is outside the timed region, the reference path inside
                layer = cpu_layer(li)
                t0 = time.perf_counter()
                nb += cpu_quantize_layer(layer, kind)
                dt += time.perf_counter() - t0
                del layer
            line["cpu_baseline"] = {"value": nb / dt / 1e9, "unit": UNIT, "cores": torch.get_num_threads() if kind == "reference" else 1,
                                    "kind": kind, "seconds": dt,
                                    "sample": f"{n_layers} of 36 Qwen3-4B decoder layers ({nb / 1e9:.2f} GB bf16) through "
                                              + ("live compressed-tensors (observer amin/amax -> calculate_qparams -> Compressor.compress)"
                                                 if kind == "reference" else "the C oracle")}
            if kind == "reference":
                # the same compressed-tensors eager ops on this GPU (SURVEY.md §8d "reference on this box"): like-for-like hardware
                try:
                    from oracle import ct_live as LCT

                    layer = [(n, f, w.to(dev)) for n, f, w in cpu_layer(0)]
                    for rep in range(2):  # first pass warms cuBLAS / allocator
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        nbc = 0
                        for _, fmt_name, w in layer:
                            fmt, a = LCT.format_args(fmt_name)
                            LCT.compress(w, fmt, a)
                            nbc += w.numel() * 2
                        torch.cuda.synchronize()
                        dtc = time.perf_counter() - t0
                    line["ct_eager_cuda"] = {"value": nbc / dtc / 1e9, "unit": UNIT,
                                             "sample": "one decoder layer through live compressed-tensors with the tensors on this GPU (eager ATen kernels)"}
                except Exception as e:  # noqa: BLE001 -- informational leg only
                    line["ct_eager_cuda"] = {"unavailable": str(e)[:200]}
        # compact per-leg summary + parity flags as the LAST keys (inside the tail the driver keeps of this line)
        def _vf(d, frac_key="frac"):
            return None if d is None else {"v": round(d["value"], 2), "u": d["unit"], "frac": round(d["roofline"][frac_key], 3) if "roofline" in d else None,
                                           "scaling": d.get("scaling")}
        legs = {"headline": {"v": round(value, 1), "u": UNIT, "frac": round(roofline["frac"], 3), "scaling": "weak"},
                "e2e": {"v": round(e2e_v, 2), "u": UNIT}}
        if strong is not None:
            legs["headline_strong"] = _vf(strong)
        if glm is not None:
            for k, v in glm.items():
                legs["glm_" + k] = _vf(v)
        if awq_line is not None:
            legs["awq"] = _vf(awq_line)
            legs["awq"]["frac_of_burst"] = round(awq_line["roofline"]["frac_of_burst"], 3) if awq_line["roofline"].get("frac_of_burst") else None
        if moe_nvfp4 is not None:
            legs["moe_nvfp4"] = _vf(moe_nvfp4)
        if moe_awq is not None:
            legs["moe_awq_experts"] = _vf(moe_awq)
            if "layer_mapping" in moe_awq:
                legs["moe_awq_layer_mapping"] = _vf(moe_awq["layer_mapping"])
                legs["moe_awq_layer_mapping"]["ms"] = round(moe_awq["layer_mapping"]["ms"], 2)
        line["parity"] = parity
        line["legs"] = legs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
