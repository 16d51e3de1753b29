"""Mixed-precision recombination: merge a base checkpoint (e.g. the vendor's FP8 weights) with the quantized tensors of a
compressed-tensors checkpoint (e.g. the AWQ W4A16 experts this package produced) into one hybrid checkpoint.

Reference behaviour: REF:scripts/recombine_weights_MiniMax-M2.1.py:75-297 --
  * ``*_proj.weight_scale_inv`` of the base is kept under the compressed-tensors name ``*_proj.weight_scale``; every other
    ``*_scale_inv`` is dropped (its weight is replaced);
  * every MoE expert weight ``...block_sparse_moe.experts.N.(w1|w2|w3).weight`` is replaced by the overlay's pack-quantized
    tensors (``weight_packed``, ``weight_scale``, ``weight_shape``, ``weight_zero_point``, ``weight_g_idx``, those that exist);
  * smoothing layers (``post_attention_layernorm.weight``) come from the overlay (AWQ rescaled them), falling back to the base;
  * everything else is copied from the base; the index is rebuilt and ``config.json`` gets a ``mixed-precision``
    ``quantization_config`` with one group per format.

This is IO only -- no arithmetic -- so nothing touches the GPU (the reference stages every shard on ``cuda:0``): tensors are
copied byte for byte between the safetensors containers (``model_free.read_header`` / ``build_header``), shard by shard, with
``pread`` / ``pwrite`` at pre-computed offsets.  The rules are data (regex lists), the MiniMax-M2.1 preset is one instance.
"""
from __future__ import annotations

import json
import os
import re
import shutil
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from .model_free import build_header, read_header

PACK_SUFFIXES = ("weight_packed", "weight_scale", "weight_shape", "weight_zero_point", "weight_g_idx")


@dataclass
class Rules:
    replace_with_packed: List[str] = field(default_factory=list)   # base ``<m>.weight`` -> the overlay's pack-quantized tensors of <m>
    take_from_overlay: List[str] = field(default_factory=list)     # same name, overlay's bytes when it has the tensor
    rename: List[Tuple[str, str]] = field(default_factory=list)    # (regex, replacement) applied to base tensor names
    drop: List[str] = field(default_factory=list)                  # base tensors left out (checked after ``rename``)
    shard_name: Optional[Callable[[str], str]] = None              # output shard name from the base shard name


MINIMAX_M21 = Rules(
    replace_with_packed=[r".*block_sparse_moe\.experts\.\d+\.(w1|w2|w3)\.weight$"],
    take_from_overlay=[r".*post_attention_layernorm.*weight$"],
    rename=[(r"(_proj\.weight)_scale_inv$", r"\1_scale")],
    drop=[r".*_scale_inv$"],
    shard_name=lambda f: f.replace("00130", "00125"),
)


def minimax_m21_quantization_config(ignore: List[str]) -> dict:
    """The hybrid config of the reference: FP8 128x128 block for every Linear, W4A16 g32 symmetric for the experts."""
    return {
        "quant_method": "compressed-tensors", "format": "mixed-precision", "quantization_status": "compressed",
        "config_groups": {
            "group_0": {"targets": ["Linear"], "format": "float-quantized",
                        "weights": {"type": "float", "num_bits": 8, "strategy": "block", "block_structure": [128, 128], "symmetric": True,
                                    "dynamic": False},
                        "input_activations": {"type": "float", "num_bits": 8, "strategy": "token", "symmetric": True, "dynamic": True}},
            "group_1": {"targets": ["Linear", r"re:.*block_sparse_moe\.experts\.\d+\.(w1|w2|w3)$"], "format": "pack-quantized",
                        "input_activations": None, "output_activations": None,
                        "weights": {"actorder": None, "block_structure": None, "dynamic": False, "group_size": 32, "num_bits": 4,
                                    "observer": "minmax", "observer_kwargs": {}, "strategy": "group", "symmetric": True, "type": "int"}},
        },
        "ignore": list(ignore), "kv_cache_scheme": None, "global_compression_ratio": None, "sparsity_config": {}, "transform_config": {},
    }


def _natural(s: str):
    return [int(t) if i & 1 else t.casefold() for i, t in enumerate(re.split(r"(\d+)", s))]


def _index(path: str) -> Dict[str, str]:
    p = os.path.join(path, "model.safetensors.index.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["weight_map"]
    out = {}
    for f in sorted(os.listdir(path)):
        if f.endswith(".safetensors"):
            for name in read_header(os.path.join(path, f))[0]:
                out[name] = f
    return out


def recombine(base_dir: str, overlay_dir: str, out_dir: str, rules: Rules = MINIMAX_M21, quantization_config: Optional[dict] = None,
              dry_run: bool = False) -> dict:
    """Merge ``base_dir`` and ``overlay_dir`` into ``out_dir`` under ``rules``; returns the statistics the reference prints
    (packed replacements, smoothing layers replaced, scale_inv renamed / skipped, total tensors, total bytes)."""
    base_map, over_map = _index(base_dir), _index(overlay_dir)
    shards = sorted(set(base_map.values()), key=_natural)
    over_hdr: Dict[str, Tuple[dict, int]] = {}

    def overlay_entry(name: str):
        f = over_map.get(name)
        if f is None:
            return None
        if f not in over_hdr:
            h, d0, _ = read_header(os.path.join(overlay_dir, f))
            over_hdr[f] = (h, d0)
        h, d0 = over_hdr[f]
        info = h[name]
        return f, info, d0 + info["data_offsets"][0], info["data_offsets"][1] - info["data_offsets"][0]

    rx = lambda pats: [re.compile(p) for p in pats]
    r_packed, r_overlay, r_drop = rx(rules.replace_with_packed), rx(rules.take_from_overlay), rx(rules.drop)
    r_rename = [(re.compile(a), b) for a, b in rules.rename]
    stats = {"pack_quantized_replaced": 0, "smoothing_layers_replaced": 0, "scale_inv_copied": 0, "scale_inv_skipped": 0, "total_tensors": 0,
             "total_size": 0, "files": 0}
    weight_map: Dict[str, str] = {}
    if not dry_run:
        os.makedirs(out_dir, exist_ok=True)
    for shard in shards:
        hdr, data0, meta = read_header(os.path.join(base_dir, shard))
        out_name = rules.shard_name(shard) if rules.shard_name else shard
        plan = []  # (out tensor name, dtype, shape, source dir, source file, absolute offset, nbytes)
        for name, info in sorted(hdr.items(), key=lambda kv: kv[1]["data_offsets"][0]):
            src = (base_dir, shard, data0 + info["data_offsets"][0], info["data_offsets"][1] - info["data_offsets"][0])
            new = name
            renamed = False
            for r, rep in r_rename:
                new2 = r.sub(rep, new)
                if new2 != new:
                    new, renamed = new2, True
            if renamed:
                plan.append((new, info["dtype"], info["shape"]) + src)
                stats["scale_inv_copied"] += 1
            elif any(r.match(name) for r in r_drop):
                stats["scale_inv_skipped"] += 1
            elif any(r.match(name) for r in r_packed):
                base = name[:-len(".weight")]
                got = 0
                for suf in PACK_SUFFIXES:
                    e = overlay_entry(f"{base}.{suf}")
                    if e is not None:
                        f, oi, off, nb = e
                        plan.append((f"{base}.{suf}", oi["dtype"], oi["shape"], overlay_dir, f, off, nb))
                        got += 1
                if got:
                    stats["pack_quantized_replaced"] += 1
            elif any(r.match(name) for r in r_overlay):
                e = overlay_entry(name)
                if e is not None:
                    f, oi, off, nb = e
                    plan.append((name, oi["dtype"], oi["shape"], overlay_dir, f, off, nb))
                    stats["smoothing_layers_replaced"] += 1
                else:
                    plan.append((name, info["dtype"], info["shape"]) + src)
            else:
                plan.append((name, info["dtype"], info["shape"]) + src)
        stats["total_tensors"] += len(plan)
        stats["total_size"] += sum(p[6] for p in plan)
        stats["files"] += 1
        for p in plan:
            weight_map[p[0]] = out_name
        if dry_run:
            continue
        head, offsets = build_header([(p[0], p[1], p[2]) for p in plan], {**meta, "format": "pt"})
        fout = os.open(os.path.join(out_dir, out_name), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        fds: Dict[str, int] = {}
        try:
            os.pwrite(fout, head, 0)
            for name, _, _, sdir, sfile, off, nb in plan:
                key = os.path.join(sdir, sfile)
                if key not in fds:
                    fds[key] = os.open(key, os.O_RDONLY)
                done = 0
                while done < nb:
                    blk = os.pread(fds[key], min(64 << 20, nb - done), off + done)
                    if not blk:
                        raise IOError(f"short read of {name} from {key}")
                    os.pwrite(fout, blk, offsets[name][0] + done)
                    done += len(blk)
        finally:
            os.close(fout)
            for fd in fds.values():
                os.close(fd)
    if not dry_run:
        with open(os.path.join(out_dir, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {"total_size": stats["total_size"]}, "weight_map": dict(sorted(weight_map.items(), key=lambda kv: _natural(kv[0])))},
                      f, indent=2)
        cfg = {}
        if os.path.exists(os.path.join(base_dir, "config.json")):
            with open(os.path.join(base_dir, "config.json")) as f:
                cfg = json.load(f)
        if quantization_config is None:
            ignore = []
            ocfg = os.path.join(overlay_dir, "config.json")
            if os.path.exists(ocfg):
                with open(ocfg) as f:
                    ignore = json.load(f).get("quantization_config", {}).get("ignore", [])
            quantization_config = minimax_m21_quantization_config(ignore)
        cfg["quantization_config"] = quantization_config
        with open(os.path.join(out_dir, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2)
        for f in os.listdir(base_dir):
            p = os.path.join(base_dir, f)
            if os.path.isfile(p) and not f.endswith(".safetensors") and f not in ("config.json", "model.safetensors.index.json"):
                shutil.copy2(p, os.path.join(out_dir, f))
    return stats
