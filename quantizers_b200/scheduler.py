"""Per-layer / per-expert work scheduler for 1..8 B200s (one process per GPU).

The reference runs everything in one process and visits Linear modules one by one
(/root/reference/scripts/do_oneshot.py:179-197; SURVEY.md §2.4 "no parallelism").  Here the units of work --
decoder layers, and inside a MoE layer the experts -- are partitioned into contiguous ranges per rank; every
weight is independent, so RTN quantize+pack needs no data-path collective.  Only observer
statistics of token-sharded calibration batches are exchanged: MIN/MAX of activation min/max, SUM of |x| sums
and of the AWQ loss accumulators, all a few KB, fused into one all-reduce per reduction op.

Weights of one shape class live stacked in one HBM arena ``[units, rows, cols]`` so a whole class is one
kernel launch (grid.z = units); siblings that must share an NVFP4 global scale (q/k/v, gate/up; LLMC
update_fused_layer_weight_global_scales) are kept on the same rank by construction (same unit index).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch


# ----------------------------------------------------------------------------- quantization presets
class SchemeArgs:
    """Attribute-compatible stand-in for compressed_tensors QuantizationArgs (quant_args.py:157-408); the ops accept
    either.  Presets mirror CT:quantization/quant_scheme.py:143-428 and the reference recipes."""

    def __init__(self, num_bits, type, symmetric, strategy, group_size=None, block_structure=None):
        self.num_bits, self.type, self.symmetric, self.strategy = num_bits, type, symmetric, strategy
        self.group_size, self.block_structure = group_size, block_structure
        self.zp_dtype = torch.int8 if type == "int" else torch.float8_e4m3fn
        self.observer = "memoryless_minmax"

    @property
    def format(self) -> str:
        """Compression format of a WEIGHT-ONLY scheme with these args (CT:compressors/format.py priority); schemes with input
        activations go through ``recipe.infer_format``.  All of naive- / int- / float-quantized share one compress arithmetic."""
        if self.type == "int":
            return "pack-quantized"
        return "naive-quantized" if self.num_bits == 8 else "nvfp4-pack-quantized"

    def bytes_per_element(self, elem_size: int = 2) -> float:
        """ALGORITHMIC HBM bytes per weight element of the fused compress."""
        code = self.num_bits / 8.0
        if self.strategy in ("group", "tensor_group"):
            scale = (1.0 if self.type == "float" and self.num_bits == 4 else elem_size) / self.group_size
            zp = 0.0 if self.symmetric else (self.num_bits / 8.0) / self.group_size
        else:
            scale, zp = 0.0, 0.0
        return elem_size + code + scale + zp


PRESETS: Dict[str, SchemeArgs] = {
    "W4A16": SchemeArgs(4, "int", True, "group", 128),                    # CT quant_scheme.py:286-295
    "W4A16_ASYM": SchemeArgs(4, "int", False, "group", 128),              # CT quant_scheme.py:298-307 (BASELINE.json)
    "INT4_G32_SYM": SchemeArgs(4, "int", True, "group", 32),              # REF:configs/recipes/recipe_awq_w4a16.yaml:17-22
    "FP8_BLOCK": SchemeArgs(8, "float", True, "block", None, [128, 128]),  # REF:scripts/quant_GLM-4.7-Flash-FP8.py:14
    "FP8_CHANNEL": SchemeArgs(8, "float", True, "channel"),               # FP8_DYNAMIC weights, CT quant_scheme.py:367-382
    "FP8_G32": SchemeArgs(8, "float", True, "group", 32),                 # REF:configs/recipes/recipe_Minimax-M2.1-AWQ-MixedPrec.yaml:26-35
    "NVFP4": SchemeArgs(4, "float", True, "tensor_group", 16),            # REF:configs/recipes/recipe_MoE_RTN_NVFP4.yaml:17-21
}


# ----------------------------------------------------------------------------- model shape tables
@dataclass
class MatrixSpec:
    name: str
    rows: int
    cols: int
    preset: str
    per_unit: int = 1          # how many such matrices per unit (e.g. k and v)
    fuse_group: Optional[str] = None  # NVFP4 siblings sharing min(global_scale)


@dataclass
class ModelSpec:
    name: str
    units: int                 # decoder layers (dense) or layers*experts (MoE expert shards)
    unit_kind: str             # "layer" | "expert"
    matrices: List[MatrixSpec] = field(default_factory=list)

    def unit_elements(self) -> int:
        return sum(m.rows * m.cols * m.per_unit for m in self.matrices)

    def unit_bytes(self, elem_size: int = 2) -> int:
        return self.unit_elements() * elem_size


def qwen3_4b(attn: str = "FP8_BLOCK", mlp: str = "W4A16_ASYM", layers: int = 36) -> ModelSpec:
    """BASELINE.json configs[1]: Qwen3-4B mixed FP8 (attn) / INT4 g128 (MLP) RTN (test-quantize_qwen3-4b-mixed-fp8-int4.yaml)."""
    return ModelSpec("qwen3-4b", layers, "layer", [
        MatrixSpec("q_proj", 4096, 2560, attn, 1, "qkv"), MatrixSpec("kv_proj", 1024, 2560, attn, 2, "qkv"),
        MatrixSpec("o_proj", 2560, 4096, attn), MatrixSpec("gate_up_proj", 9728, 2560, mlp, 2, "gate_up"),
        MatrixSpec("down_proj", 2560, 9728, mlp)])


def qwen3_30b_a3b(preset: str = "NVFP4", layers: int = 48, experts: int = 128) -> ModelSpec:
    """BASELINE.json configs[3]: per expert gate, up [768,2048], down [2048,768]."""
    return ModelSpec("qwen3-30b-a3b", layers * experts, "expert", [
        MatrixSpec("gate_up_proj", 768, 2048, preset, 2, "gate_up"), MatrixSpec("down_proj", 2048, 768, preset)])


def minimax_m21_experts(preset: str = "INT4_G32_SYM", layers: int = 62, experts: int = 256) -> ModelSpec:
    """BASELINE.json configs[4]: per expert w1, w3 [1536,3072], w2 [3072,1536]."""
    return ModelSpec("minimax-m2.1", layers * experts, "expert", [
        MatrixSpec("w1_w3", 1536, 3072, preset, 2, "gate_up"), MatrixSpec("w2", 3072, 1536, preset)])


def glm47_flash(preset: str = "FP8_BLOCK", units: int = 64) -> ModelSpec:
    """BASELINE.json configs[2]: dims unverified (no network) -> synthetic shapes incl. a non-multiple-of-128 row count."""
    return ModelSpec("glm-4.7-flash", units, "expert", [
        MatrixSpec("gate_up_proj", 1536, 2048, preset, 2), MatrixSpec("down_proj", 2048, 1536, preset),
        MatrixSpec("dense_up", 10240, 2048, preset), MatrixSpec("dense_ragged", 2560 + 64, 2048, preset)])


# ----------------------------------------------------------------------------- partitioning
def partition(n_units: int, world_size: int, rank: int) -> range:
    """Contiguous balanced ranges; the first (n_units % world_size) ranks take one extra unit."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, extra = divmod(n_units, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def partition_balanced(spec: ModelSpec, n_units: int, world_size: int) -> List[Dict[str, List[int]]]:
    """Strong-scaling partition at (matrix class, unit) granularity, balanced by elements: per rank {class name: [unit ids]}.

    ``partition`` hands out whole units; 36 layers over 8 ranks is 5 / 4 layers, i.e. at best 7.2x.  The classes of a unit are
    independent (RTN: no data-path exchange), so they can go to different ranks: items are sorted by size and each goes to the
    least-loaded rank (LPT; deterministic, identical on every rank).  Classes whose NVFP4 siblings share min(global_scale) across
    classes (``fuse_group``) stay on one rank."""
    if world_size < 1:
        raise ValueError(f"bad world_size {world_size}")
    groups: Dict[str, List[MatrixSpec]] = {}
    for m in spec.matrices:
        a = PRESETS[m.preset]
        tied = m.fuse_group is not None and a.type == "float" and a.num_bits == 4
        groups.setdefault(m.fuse_group if tied else m.name, []).append(m)
    items = []
    for gname, members in groups.items():
        wgt = sum(m.rows * m.cols * m.per_unit for m in members)
        items.extend((wgt, gname, u) for u in range(n_units))
    items.sort(key=lambda t: (-t[0], t[1], t[2]))
    load = [0] * world_size
    assign: List[Dict[str, List[int]]] = [{} for _ in range(world_size)]
    for wgt, gname, u in items:
        r = min(range(world_size), key=lambda i: (load[i], i))
        load[r] += wgt
        for m in groups[gname]:
            assign[r].setdefault(m.name, []).append(u)
    for a in assign:
        for units in a.values():
            units.sort()
    return assign


def build_arena_classes(spec: ModelSpec, units_by_class: Dict[str, Sequence[int]], device, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """``build_arena`` with a separate unit list per matrix class (``partition_balanced``); classes without units are left out --
    ``alloc_outputs`` / ``quantize_arena`` skip them."""
    arena = {}
    for mi, m in enumerate(spec.matrices):
        units = list(units_by_class.get(m.name, ()))
        if not units:
            continue
        stacks = [synth_stack(units, m.rows, m.cols, mi * 8 + j, device, dtype) for j in range(m.per_unit)]
        arena[m.name] = torch.stack(stacks, dim=1).reshape(len(units) * m.per_unit, m.rows, m.cols) if m.per_unit > 1 else stacks[0]
    return arena


def owner_of(unit: int, n_units: int, world_size: int) -> int:
    base, extra = divmod(n_units, world_size)
    cut = extra * (base + 1)
    return unit // (base + 1) if unit < cut else extra + (unit - cut) // max(base, 1)


# ----------------------------------------------------------------------------- synthetic arenas
def synth_stack(units: Sequence[int], rows: int, cols: int, matrix_idx: int, device, dtype=torch.bfloat16,
                outliers: bool = True) -> torch.Tensor:
    """W = randn * 0.02 seeded by 1234 + unit*1000 + matrix_idx; 0.1 % of the columns multiplied by 20."""
    out = torch.empty((len(units), rows, cols), dtype=dtype, device=device)
    for i, u in enumerate(units):
        g = torch.Generator(device=device).manual_seed(1234 + int(u) * 1000 + matrix_idx)
        w = torch.randn(rows, cols, generator=g, device=device, dtype=torch.float32) * 0.02
        if outliers:
            step = max(cols // max(cols // 1000, 1), 1)
            w[:, ::step] *= 20.0
        out[i] = w.to(dtype)
    return out


def build_arena(spec: ModelSpec, units: Sequence[int], device, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    arena = {}
    for mi, m in enumerate(spec.matrices):
        stacks = [synth_stack(units, m.rows, m.cols, mi * 8 + j, device, dtype) for j in range(m.per_unit)]
        arena[m.name] = torch.stack(stacks, dim=1).reshape(len(units) * m.per_unit, m.rows, m.cols) if m.per_unit > 1 else stacks[0]
    return arena


def synth_awq_layer(layer_idx: int, tokens: int, device, hidden: int = 2560, inter: int = 9728, n_heads: int = 32, n_kv: int = 8,
                    head_dim: int = 128, dtype=torch.bfloat16):
    """One Qwen3-4B-shaped decoder layer + the balance-layer inputs of its AWQ mappings: weights
    randn*0.02 (seed 1234 + layer*1000 + matrix), activations randn * (1 + 3 rand(K)) (seed 4321 + layer) so that the
    per-channel means x_mean spread; down_proj's input is the layer's own silu(gate x) * up x."""
    shapes = {"q": (n_heads * head_dim, hidden), "k": (n_kv * head_dim, hidden), "v": (n_kv * head_dim, hidden),
              "o": (hidden, n_heads * head_dim), "gate": (inter, hidden), "up": (inter, hidden), "down": (hidden, inter)}
    w = {}
    for mi, (name, (r, c)) in enumerate(shapes.items()):
        w[name] = synth_stack([layer_idx], r, c, mi, device, dtype)[0]
    g = torch.Generator(device=device).manual_seed(4321 + layer_idx)
    for name, n in (("input_layernorm", hidden), ("post_attention_layernorm", hidden), ("q_norm", head_dim), ("k_norm", head_dim)):
        w[name] = (1 + 0.1 * torch.randn(n, generator=g, device=device)).to(dtype)
    acts = {}
    for name in ("attn_in", "mlp_in"):
        spread = 1 + 3 * torch.rand(hidden, generator=g, device=device)
        acts[name] = (torch.randn(tokens, hidden, generator=g, device=device) * spread).to(dtype)
    F = torch.nn.functional
    acts["down_in"] = torch.empty(tokens, inter, dtype=dtype, device=device)
    for t0 in range(0, tokens, 8192):
        xc = acts["mlp_in"][t0:t0 + 8192]
        acts["down_in"][t0:t0 + 8192] = F.silu(F.linear(xc, w["gate"])) * F.linear(xc, w["up"])
    return w, acts


def synth_moe_awq_experts(layer_idx: int, experts: Sequence[int], tokens: int, device, hidden: int = 3072, inter: int = 1536,
                          dtype=torch.bfloat16):
    """The ``w3 -> w2`` mappings of MiniMax-M2.1-shaped experts: per expert w1, w3 [inter, hidden],
    w2 [hidden, inter] (seed 1234 + (layer * 256 + expert) * 1000 + matrix) and the balance-layer input
    ``silu(w1 x) * (w3 x)`` [tokens, inter] of the layer's calibration activations x (seed 4321 + layer; every expert sees
    all tokens, moe_calibrate_all_experts).  Returns (w1, w3, w2 stacked [E, ...], list of inputs)."""
    units = [layer_idx * 256 + int(e) for e in experts]
    w1 = synth_stack(units, inter, hidden, 0, device, dtype)
    w3 = synth_stack(units, inter, hidden, 1, device, dtype)
    w2 = synth_stack(units, hidden, inter, 2, device, dtype)
    g = torch.Generator(device=device).manual_seed(4321 + layer_idx)
    spread = 1 + 3 * torch.rand(hidden, generator=g, device=device)
    x = (torch.randn(tokens, hidden, generator=g, device=device) * spread).to(dtype)
    F = torch.nn.functional
    xs = []
    for e in range(len(units)):
        h = torch.empty(tokens, inter, dtype=dtype, device=device)
        for t0 in range(0, tokens, 8192):
            xc = x[t0:t0 + 8192]
            h[t0:t0 + 8192] = F.silu(F.linear(xc, w1[e])) * F.linear(xc, w3[e])
        xs.append(h)
    return w1, w3, w2, xs


# ----------------------------------------------------------------------------- RTN quantize + pack of a shard
def _nvfp4_span(spec: ModelSpec, m: MatrixSpec) -> int:
    """Stacked siblings that share min(global_scale) inside ONE stack (gate/up: per_unit = 2), else 1."""
    a = PRESETS[m.preset]
    if not (a.type == "float" and a.num_bits == 4):
        return 1
    members = [x for x in spec.matrices if (x.fuse_group or x.name) == (m.fuse_group or m.name)
               and PRESETS[x.preset].type == "float" and PRESETS[x.preset].num_bits == 4]
    return m.per_unit if len(members) == 1 else 1


def alloc_outputs(spec: ModelSpec, arena: Dict[str, torch.Tensor]) -> Dict[str, dict]:
    """Caller-owned output + workspace buffers for :func:`quantize_arena` (one set per shape class, reused for every pass): with
    them a pass allocates nothing and never synchronises with the host."""
    from . import ops

    return {m.name: ops.compress_outputs(arena[m.name].shape, PRESETS[m.preset], arena[m.name].dtype, arena[m.name].device,
                                         fuse_span=_nvfp4_span(spec, m)) for m in spec.matrices if m.name in arena}


_CLASS_STREAMS: Dict[int, list] = {}


def _class_streams(device, n: int) -> list:
    key = torch.device(device).index or 0
    pool = _CLASS_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device))
    return pool[:n]


def quantize_arena(spec: ModelSpec, arena: Dict[str, torch.Tensor], timings: Optional[list] = None,
                   out: Optional[Dict[str, dict]] = None, concurrent: bool = False) -> Dict[str, dict]:
    """Fused observe -> qparams -> quantize -> pack for every stacked weight class of this rank's shard.

    ``out``: buffers from :func:`alloc_outputs`; results are written in place (no allocation, no host sync).

    NVFP4: siblings in one ``fuse_group`` of a unit share min(global_scale) (LLMC
    update_fused_layer_weight_global_scales); the per-tensor scales come from one batched reduction per class.
    ``timings``: when a list is given, (class name, preset, elements, start_event, end_event) is appended per launch.
    ``concurrent``: the classes are independent, so each is launched on its own side stream (fork / join on events around the pass;
    capturable in a CUDA graph): when a rank's stacks are small (strong scaling at N = 8: 20-200 us per launch) one launch's tail
    overlaps the next one's ramp-up.  Needs ``out`` (no allocation on side streams) and no ``timings``.
    """
    from . import ops

    res = {}
    fused_gs: Dict[str, torch.Tensor] = {}
    nv = [m for m in spec.matrices if m.name in arena and PRESETS[m.preset].type == "float" and PRESETS[m.preset].num_bits == 4]
    groups = {}
    for m in nv:
        groups.setdefault(m.fuse_group or m.name, []).append(m)
    span_of: Dict[str, int] = {}
    for gname, members in list(groups.items()):
        if len(members) == 1:
            # all siblings of the group sit next to each other in ONE stack (gate/up: per_unit = 2): the fused kernel computes
            # |max| -> min(global_scale) -> codes in a single launch with one HBM read
            span_of[members[0].name] = members[0].per_unit
            del groups[gname]
    for gname, members in groups.items():
        per_unit_min = None
        for m in members:
            w = arena[m.name]
            gs = ops.weight_global_scales(w).reshape(-1, m.per_unit).amin(dim=1)  # [units]
            per_unit_min = gs if per_unit_min is None else torch.minimum(per_unit_min, gs)
        for m in members:
            fused_gs[m.name] = per_unit_min.repeat_interleave(m.per_unit).contiguous()
    present = [m for m in spec.matrices if m.name in arena]
    if concurrent and out is not None and timings is None and len(present) > 1:
        dev = arena[present[0].name].device
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        for m, st in zip(present, _class_streams(dev, len(present))):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                res[m.name] = ops.compress_weight(arena[m.name], PRESETS[m.preset], global_scale=fused_gs.get(m.name),
                                                  fuse_span=span_of.get(m.name, 1), out=out[m.name])
            join = torch.cuda.Event()
            join.record(st)
            cur.wait_event(join)
        return res
    for m in spec.matrices:
        if m.name not in arena:        # a class this rank holds no unit of (partition_balanced)
            continue
        w = arena[m.name]
        args = PRESETS[m.preset]
        ev = None
        if timings is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        res[m.name] = ops.compress_weight(w, args, global_scale=fused_gs.get(m.name), fuse_span=span_of.get(m.name, 1),
                                          out=None if out is None else out[m.name])
        if ev is not None:
            ev[1].record()
            timings.append((m.name, m.preset, w.numel(), ev[0], ev[1]))
    return res


def launches_per_step(spec: ModelSpec) -> int:
    """Kernel launches of one quantize_arena pass (our kernels only)."""
    n = 0
    for m in spec.matrices:
        a = PRESETS[m.preset]
        if a.type == "float" and a.num_bits == 4:
            n += 1  # fused |max| -> global scale -> compress kernel (siblings in one stack)
        elif a.strategy == "tensor":
            n += 3
        elif a.type == "int" and not a.symmetric and a.strategy == "group":
            n += 2  # fused compress + zero-point row-pack kernel
        else:
            n += 1
    return n


# ----------------------------------------------------------------------------- statistic exchange (token-sharded calibration)
def allreduce_stats(mins: Optional[torch.Tensor] = None, maxs: Optional[torch.Tensor] = None, sums: Optional[torch.Tensor] = None,
                    group=None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """One all-reduce per reduction op over flat fp32 buffers (MIN / MAX are order independent => bit-identical for
    any world size; SUM differs in the last bits => AWQ argmin uses the first-minimum tie policy)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return mins, maxs, sums
    if mins is not None:
        dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
    if maxs is not None:
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    if sums is not None:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return mins, maxs, sums
