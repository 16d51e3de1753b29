"""``model_free_ptq``: data-free weight quantization of a safetensors checkpoint, shard by shard, without building the model.

Reference call site: REF:scripts/quant_GLM-4.7-Flash-FP8.py:11-24

    model_free_ptq(model_stub, save_directory, scheme="FP8_BLOCK", ignore=[...], max_workers=16, device="cuda:0")

(llmcompressor ``entrypoints/model_free``; restated in SURVEY.md §3.3).  Every 2-D ``<module>.weight`` whose module name is not
ignored is run through the fused observer -> qparams -> quantize -> pack kernel and written in the compressed-tensors layout
(``weight`` e4m3 + ``weight_scale`` for float-quantized, ``weight_packed`` / ``weight_scale`` / ``weight_zero_point`` /
``weight_shape`` for pack-quantized, ``weight_packed`` / ``weight_scale`` / ``weight_global_scale`` for nvfp4); everything else is
copied through byte for byte; ``model.safetensors.index.json`` and ``config.json`` (``quantization_config``) are rewritten.

At multi-TB/s kernels the job is IO: each worker thread owns one ``b200q_pipeline`` (three device slots, separate H2D / run / D2H
streams) and a ring of pinned staging buffers; tensors are read straight from the file into pinned memory (``readinto``, GIL
released), cross PCIe once in each direction, and leave with ``os.pwrite`` at their pre-computed offset in the output shard (the
output header is known from the shapes alone, so nothing is buffered).  Shards are independent: with ``world_size > 1`` rank r
takes the shards ``i % world_size == r`` and rank 0 writes the index / config -- no collective.

The safetensors container is parsed here (8-byte little-endian header length, JSON header, raw little-endian data) so that the
tensors can be addressed as plain host pointers for the C ABI.
"""
from __future__ import annotations

import ctypes
import json
import os
import shutil
import struct
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import numa, ops
from .recipe import match_name, preset_args, preset_input_args
from .scheduler import SchemeArgs

_ST_DTYPES = {"BF16": (torch.bfloat16, 2), "F16": (torch.float16, 2), "F32": (torch.float32, 4), "F64": (torch.float64, 8),
              "I64": (torch.int64, 8), "I32": (torch.int32, 4), "I16": (torch.int16, 2), "I8": (torch.int8, 1), "U8": (torch.uint8, 1),
              "BOOL": (torch.bool, 1), "F8_E4M3": (torch.float8_e4m3fn, 1), "F8_E5M2": (torch.float8_e5m2, 1)}
_QUANTIZABLE = ("BF16", "F16", "F32")
MAX_GPU_WORKERS = 4


# ----------------------------------------------------------------------------- safetensors container
def read_header(path: str) -> Tuple[Dict[str, dict], int, dict]:
    """-> ({tensor name: {dtype, shape, data_offsets}}, byte offset of the data section, __metadata__)."""
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        if n > 512 << 20:
            raise ValueError(f"{path}: implausible safetensors header length {n}")
        hdr = json.loads(f.read(n).decode("utf-8"))
    meta = hdr.pop("__metadata__", {}) or {}
    return hdr, 8 + n, meta


def _numel(shape: Sequence[int]) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


def build_header(entries: List[Tuple[str, str, Sequence[int]]], metadata: Optional[dict] = None) -> Tuple[bytes, Dict[str, Tuple[int, int]]]:
    """entries: (name, safetensors dtype, shape) in file order -> (header bytes incl. the length prefix, {name: (offset, nbytes)}
    relative to the start of the file)."""
    hdr: Dict[str, dict] = {}
    if metadata:
        hdr["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    off = 0
    rel: Dict[str, Tuple[int, int]] = {}
    for name, dt, shape in entries:
        nbytes = _numel(shape) * _ST_DTYPES[dt][1]
        hdr[name] = {"dtype": dt, "shape": [int(d) for d in shape], "data_offsets": [off, off + nbytes]}
        rel[name] = (off, nbytes)
        off += nbytes
    raw = json.dumps(hdr, separators=(",", ":")).encode("utf-8")
    raw += b" " * ((8 - len(raw) % 8) % 8)  # keep the data section 8-byte aligned, as the reference writer does
    head = struct.pack("<Q", len(raw)) + raw
    return head, {k: (len(head) + o, n) for k, (o, n) in rel.items()}


# ----------------------------------------------------------------------------- plan
def _scheme_args(scheme) -> SchemeArgs:
    if isinstance(scheme, str):
        a = preset_args(scheme)
        if a is None:
            raise ValueError("scheme UNQUANTIZED leaves nothing to do")
        return a
    return scheme


def _is_ignored(module: str, ignore: Iterable[str]) -> bool:
    return any(match_name(module, i) for i in ignore)


def compressed_entries(module: str, rows: int, cols: int, dt: str, a: SchemeArgs) -> List[Tuple[str, str, Tuple[int, ...]]]:
    """The state-dict entries ``Compressor.compress`` emits for one Linear weight, in write order."""
    if a.type == "int":
        pf = 32 // a.num_bits
        g = (rows, cols // a.group_size) if a.strategy == "group" else (rows, 1)
        out = [(f"{module}.weight_packed", "I32", (rows, -(-cols // pf))), (f"{module}.weight_scale", dt, g)]
        if not a.symmetric:
            out.append((f"{module}.weight_zero_point", "I32", (-(-g[0] // pf), g[1])))
        out.append((f"{module}.weight_shape", "I64", (2,)))
        return out
    if a.num_bits == 8:
        if a.strategy == "block":
            bh, bw = a.block_structure
            g: Tuple[int, ...] = (-(-rows // bh), -(-cols // bw))
        elif a.strategy == "channel":
            g = (rows, 1)
        elif a.strategy == "group":
            g = (rows, cols // a.group_size)
        else:
            g = (1,)
        return [(f"{module}.weight", "F8_E4M3", (rows, cols)), (f"{module}.weight_scale", dt, g)]
    return [(f"{module}.weight_packed", "U8", (rows, cols // 2)), (f"{module}.weight_scale", "F8_E4M3", (rows, cols // 16)),
            (f"{module}.weight_global_scale", "F32", (1,))]


def _check_shape(module: str, rows: int, cols: int, a: SchemeArgs):
    if a.strategy in ("group", "tensor_group") and cols % a.group_size != 0:
        raise ValueError(f"{module}: tensor column shape must be divisible by the given group_size {a.group_size} but got {cols}")
    if a.type == "float" and a.num_bits == 4 and cols % 16 != 0:
        raise ValueError(f"{module}: NVFP4 needs a multiple of 16 columns, got {cols}")


# ----------------------------------------------------------------------------- worker
class _Worker:
    """One thread's resources: a C pipeline handle and pinned staging rings for inputs and outputs."""

    RING = 3

    def __init__(self, max_weight_bytes: int, device_index: int):
        self.lib = L.lib()
        self.handle = ctypes.c_void_p()
        L.check(self.lib.b200q_pipeline_create(ctypes.byref(self.handle), max_weight_bytes, device_index))
        # one pinned arena per worker (a single cudaHostAlloc: page-locking is the expensive part of start-up), carved into the rings
        sizes = [max_weight_bytes, max_weight_bytes // 2 + 256, max_weight_bytes // 16 + 4096, max_weight_bytes // 64 + 4096, 16]
        sizes = [(n + 255) // 256 * 256 for n in sizes]
        with numa.near_device(device_index):  # first touch on the GPU's own NUMA node
            self._arena = torch.empty(self.RING * sum(sizes), dtype=torch.uint8, pin_memory=True)
        off = 0
        rings = []
        for n in sizes:
            rings.append([self._arena[off + i * n: off + (i + 1) * n] for i in range(self.RING)])
            off += self.RING * n
        self.w, self.codes, self.scale, self.zp = rings[0], rings[1], rings[2], rings[3]
        self.gs = [t[:16].view(torch.float32) for t in rings[4]]
        self.pending: List[tuple] = []
        self.slot = 0

    def submit(self, fin, src_off: int, nbytes: int, rows: int, cols: int, dt: str, a: SchemeArgs, fout: int, dst: Dict[str, Tuple[int, int]],
               module: str):
        if len(self.pending) >= self.RING - 1:
            self.drain(1)
        s = self.slot
        self.slot = (self.slot + 1) % self.RING
        buf = self.w[s].numpy()[:nbytes]
        got = os.preadv(fin, [memoryview(buf)], src_off)
        if got != nbytes:
            raise IOError(f"short read for {module}.weight: {got} of {nbytes} bytes")
        sc = ops.scheme_from_args(a, _ST_DTYPES[dt][0], True)
        L.check(self.lib.b200q_pipeline_compress_host(self.handle, L.ptr(self.w[s]), 1, rows, cols, ctypes.byref(sc), L.ptr(self.codes[s]),
                                                      L.ptr(self.scale[s]), L.ptr(self.zp[s]), L.ptr(self.gs[s])))
        self.pending.append((s, fout, dst, module, rows, cols, a))

    def drain(self, keep: int = 0):
        """Write finished jobs out.  The pipeline is in order, so one sync finishes everything submitted so far."""
        if len(self.pending) <= keep:
            return
        L.check(self.lib.b200q_pipeline_sync(self.handle))
        for s, fout, dst, module, rows, cols, a in self.pending:
            for key, (off, nbytes) in dst.items():
                leaf = key[len(module) + 1:]
                if leaf in ("weight_packed", "weight"):
                    src = self.codes[s]
                elif leaf == "weight_scale":
                    src = self.scale[s]
                elif leaf == "weight_zero_point":
                    src = self.zp[s]
                elif leaf == "weight_global_scale":
                    src = self.gs[s].view(torch.uint8)
                else:  # weight_shape
                    src = torch.tensor([rows, cols], dtype=torch.int64).view(torch.uint8)
                if nbytes > src.numel():  # b200q_pipeline_compress_host refuses such jobs up front; never truncate silently
                    raise RuntimeError(f"{key}: {nbytes} bytes do not fit the {src.numel()}-byte staging slot")
                os.pwrite(fout, memoryview(src.numpy()[:nbytes]), off)
        self.pending = []

    def close(self):
        self.drain()
        self.lib.b200q_pipeline_destroy(self.handle)


# ----------------------------------------------------------------------------- entry point
def model_free_ptq(model_stub: str, save_directory: str, scheme="FP8_BLOCK", ignore: Sequence[str] = (), max_workers: int = 4,
                   device: str = "cuda:0", rank: int = 0, world_size: int = 1) -> dict:
    """Quantize every non-ignored 2-D ``*.weight`` of the safetensors checkpoint in ``model_stub`` (a local directory; there is no
    network here) into ``save_directory``.  Returns a summary {files, tensors_quantized, tensors_copied, bytes_in, bytes_out}."""
    if not os.path.isdir(model_stub):
        raise FileNotFoundError(f"{model_stub!r} is not a local checkpoint directory (hub downloads are not available)")
    if not torch.cuda.is_available():
        raise RuntimeError("model_free_ptq needs a CUDA device: the quantization hot path has no CPU fallback")
    a = _scheme_args(scheme)
    dev = torch.device(device)
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    os.makedirs(save_directory, exist_ok=True)
    shards = sorted(f for f in os.listdir(model_stub) if f.endswith(".safetensors"))
    if not shards:
        raise FileNotFoundError(f"no .safetensors files in {model_stub}")
    mine = [f for i, f in enumerate(shards) if i % world_size == rank]

    # ---- plan every shard: output entries, offsets, job list
    plans = []
    max_bytes = 1 << 20
    weight_map: Dict[str, str] = {}
    for fname in shards:
        hdr, data0, meta = read_header(os.path.join(model_stub, fname))
        order = sorted(hdr.items(), key=lambda kv: kv[1]["data_offsets"][0])
        entries, jobs, copies = [], [], []
        for name, info in order:
            module = name[:-len(".weight")] if name.endswith(".weight") else None
            shape = info["shape"]
            if (module is not None and len(shape) == 2 and info["dtype"] in _QUANTIZABLE and not _is_ignored(module, ignore)):
                rows, cols = int(shape[0]), int(shape[1])
                _check_shape(module, rows, cols, a)
                ent = compressed_entries(module, rows, cols, info["dtype"], a)
                entries += ent
                jobs.append((module, rows, cols, info["dtype"], data0 + info["data_offsets"][0], info["data_offsets"][1] - info["data_offsets"][0],
                             [e[0] for e in ent]))
                max_bytes = max(max_bytes, info["data_offsets"][1] - info["data_offsets"][0])
            else:
                entries.append((name, info["dtype"], shape))
                copies.append((name, data0 + info["data_offsets"][0], info["data_offsets"][1] - info["data_offsets"][0]))
        head, offsets = build_header(entries, {**meta, "format": "pt"})
        for e in entries:
            weight_map[e[0]] = fname
        plans.append((fname, head, offsets, jobs, copies))

    stats = {"files": 0, "tensors_quantized": 0, "tensors_copied": 0, "bytes_in": 0, "bytes_out": 0}
    lock = threading.Lock()
    local = threading.local()
    workers: List[_Worker] = []

    def worker() -> _Worker:
        w = getattr(local, "w", None)
        if w is None:
            torch.cuda.set_device(dev_index)
            w = local.w = _Worker(max_bytes, dev_index)
            with lock:
                workers.append(w)
        return w

    def run_chunk(fin: int, fout: int, offsets, jobs):
        w = worker()
        for module, rows, cols, dt, src_off, nbytes, keys in jobs:
            w.submit(fin, src_off, nbytes, rows, cols, dt, a, fout, {k: offsets[k] for k in keys}, module)
        w.drain()

    # Every worker owns pinned staging rings and a device pipeline sized for the largest tensor; page-locking that memory is the
    # dominant start-up cost and PCIe is shared, so more than MAX_GPU_WORKERS pipelines only slow the job down (measured on a
    # 5.2 GB checkpoint: 4 workers 3.6 GB/s, 8 workers 1.5 GB/s, 16 workers 0.9 GB/s).
    max_workers = min(max(1, int(max_workers)), MAX_GPU_WORKERS)
    with ThreadPoolExecutor(max_workers=max_workers) as pool:
        for fname, head, offsets, jobs, copies in plans:
            if fname not in mine:
                continue
            fin = os.open(os.path.join(model_stub, fname), os.O_RDONLY)
            fout = os.open(os.path.join(save_directory, fname), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            try:
                os.ftruncate(fout, max((o + n for o, n in offsets.values()), default=len(head)))
                os.pwrite(fout, head, 0)
                n_chunks = max(1, min(int(max_workers), len(jobs)))
                futs = [pool.submit(run_chunk, fin, fout, offsets, jobs[i::n_chunks]) for i in range(n_chunks)] if jobs else []
                for name, src_off, nbytes in copies:  # pass-through tensors: plain byte copy on this thread
                    done = 0
                    while done < nbytes:
                        blk = os.pread(fin, min(64 << 20, nbytes - done), src_off + done)
                        os.pwrite(fout, blk, offsets[name][0] + done)
                        done += len(blk)
                for f in futs:
                    f.result()
                stats["files"] += 1
                stats["tensors_quantized"] += len(jobs)
                stats["tensors_copied"] += len(copies)
                stats["bytes_in"] += os.fstat(fin).st_size
                stats["bytes_out"] += os.fstat(fout).st_size
            finally:
                os.close(fin)
                os.close(fout)
    for w in workers:
        w.close()

    if rank == 0:
        # a preset name carries its input-activation half into quantization_config (FP8_BLOCK: dynamic group-128 fp8 inputs)
        _write_sidecars(model_stub, save_directory, a, list(ignore), weight_map, shards,
                        preset_input_args(scheme) if isinstance(scheme, str) else None)
    return stats


def _write_sidecars(src: str, dst: str, a: SchemeArgs, ignore: List[str], weight_map: Dict[str, str], shards: List[str],
                    input_activations: Optional[SchemeArgs] = None):
    """index json with the new tensor names, config.json with the quantization_config block, every other small file copied."""
    for f in os.listdir(src):
        p = os.path.join(src, f)
        if f.endswith(".safetensors") or not os.path.isfile(p) or f in ("model.safetensors.index.json", "config.json"):
            continue
        shutil.copy2(p, os.path.join(dst, f))
    if len(shards) > 1 or os.path.exists(os.path.join(src, "model.safetensors.index.json")):
        total = 0
        for f in shards:
            q = os.path.join(dst, f)
            if os.path.exists(q):
                total += os.path.getsize(q)
        with open(os.path.join(dst, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {"total_size": total}, "weight_map": dict(sorted(weight_map.items()))}, f, indent=2)
    cfg_path = os.path.join(src, "config.json")
    cfg = {}
    if os.path.exists(cfg_path):
        with open(cfg_path) as f:
            cfg = json.load(f)
    from .recipe import ConfigGroup, group_to_dict

    grp = group_to_dict(ConfigGroup("group_0", ["Linear"], a, input_activations))
    cfg["quantization_config"] = {"quant_method": "compressed-tensors", "format": grp["format"], "quantization_status": "compressed",
                                  "ignore": ignore, "config_groups": {"group_0": grp}}
    with open(os.path.join(dst, "config.json"), "w") as f:
        json.dump(cfg, f, indent=2)
