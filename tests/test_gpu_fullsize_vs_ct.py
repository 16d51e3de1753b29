"""GPU: BASELINE.json-size matrices, bit-exact against the reference's OWN arithmetic run on this GPU.

The reference's quantize path is compressed-tensors (CT) eager torch ops; with the tensors on the B200 the whole
``observer -> calculate_qparams -> Compressor.compress`` chain of a full-size matrix is a few milliseconds, so every BASELINE config
shape is compared in full -- all rows, every code, scale and zero point (round-1 verdict, "What's missing" #1):

  config 2 (Qwen3-4B mixed):  FP8 128x128 on q/k/v/o shapes, INT4 g128 asym and g32 sym on gate/up/down shapes
  config 3 (GLM-4.7-Flash):   FP8 128x128 incl. the ragged [2624, 9728] / [2624, 2048], FP8 per-channel
  config 4 (Qwen3-30B-A3B):   NVFP4 expert stacks [128 x 2, 768, 2048] (gate/up share min(global_scale)) and [128, 2048, 768]
  config 5 (MiniMax-M2.1):    INT4 g32 sym expert stacks [.., 1536, 3072] / [.., 3072, 1536]

First ``test_ct_cuda_equals_ct_cpu`` pins CT-on-CUDA to CT-on-CPU (which the C oracle and the golden fixtures are pinned to) on
every format, so "equal to CT on this GPU" means "equal to the reference".
Reference anchors: CT:compressors/pack_quantized/base.py:36-77, CT:compressors/naive_quantized/base.py:34-82,
CT:compressors/nvfp4/base.py:40-72, CT:quantization/utils/helpers.py:50-137,309-338.
"""
import pytest
import torch

from oracle import ct_live as L
from tests.test_gpu_compress import Args
from tests.util import FORMATS, assert_bits_equal, synth_weight

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not L.available(), reason="compressed_tensors is not importable")]


def _w(R, C, seed, outliers=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(R, C, generator=g, device="cuda") * 0.02
    if outliers:
        w[:, ::997] *= 20
        w[R // 2, : min(C, 256)] = 0.0   # all-zero groups (eps / NaN zero-point path)
        w[R // 3, :] = -0.0              # sign-nibble path
    return w.to(torch.bfloat16)


def _same(got: dict, want: dict, what: str):
    for k, v in want.items():
        if k == "weight_shape":
            assert got[k].device.type == "cpu" and got[k].dtype == torch.int64, f"{what}: weight_shape must be a CPU int64 tensor like CT's"
            assert got[k].tolist() == v.tolist(), f"{what}: weight_shape"
            continue
        assert k in got, f"{what}: missing key {k}"
        assert got[k].dtype == v.dtype, f"{what}:{k} dtype {got[k].dtype} != {v.dtype}"
        assert_bits_equal(got[k].reshape(v.shape), v, f"{what}:{k}")
    extra = set(got) - set(want)
    assert not extra, f"{what}: unexpected keys {extra}"


@pytest.mark.parametrize("name", [n for n in FORMATS if n != "int8_g128_sym"])
def test_ct_cuda_equals_ct_cpu(name):
    """The oracle anchor: live CT gives the same bits on the CPU and on this GPU (edge cases included)."""
    fmt, a = L.format_args(name)
    w = synth_weight(200 if "block" in name else 128, 1280, torch.bfloat16, 11)
    cpu = L.compress(w, fmt, a)
    gpu = L.compress(w.cuda(), fmt, a)
    assert set(cpu) == set(gpu)
    for k in cpu:
        assert_bits_equal(gpu[k], cpu[k], f"{name}:{k}")


CASES = [
    # config 2: attention -> FP8_BLOCK, MLP -> INT4 (g128 asym per BASELINE.json, g32 sym per the reference's recipes)
    ("fp8_block", (4096, 2560)), ("fp8_block", (1024, 2560)), ("fp8_block", (2560, 4096)),
    ("int4_g128_asym", (9728, 2560)), ("int4_g128_asym", (2560, 9728)),
    ("int4_g32_sym", (9728, 2560)), ("int4_g32_sym", (2560, 9728)),
    ("int4_g128_sym", (9728, 2560)), ("int4_g32_asym", (2560, 9728)),
    # config 3: FP8 block incl. non-multiple-of-128 rows, FP8 per-channel
    ("fp8_block", (2624, 9728)), ("fp8_block", (2624, 2048)), ("fp8_block", (1536, 2048)), ("fp8_block", (10240, 2048)),
    ("fp8_channel", (2560, 4096)), ("fp8_channel", (2624, 2048)),
    # MiniMax mixed-precision AWQ recipe: FP8 g32
    ("fp8_g32", (1536, 3072)),
    # W8A8 weights (REF:scripts/quantization_multiple_modifiers.py:55: self_attn to W8A8): symmetric INT8 per channel, one-, two- and
    # four-warp row teams of channel_fast_kernel
    # CT's "FP8" preset: one static scale per weight (two-pass bf16 fast path)
    ("fp8_tensor", (4096, 2560)), ("fp8_tensor", (2624, 2048)), ("fp8_tensor", (37, 1000)),
    ("int8_channel_sym", (4096, 2560)), ("int8_channel_sym", (2560, 4096)), ("int8_channel_sym", (2560, 9728)), ("int8_channel_sym", (1000, 2056)),
]


@pytest.mark.parametrize("name,shape", CASES)
def test_fullsize_bit_exact_vs_ct_cuda(name, shape):
    from quantizers_b200 import ops

    fmt, a = L.format_args(name)
    R, C = shape
    w = _w(R, C, 1234 + R + C)
    want = L.compress(w, fmt, a)
    got = ops.compress_weight(w, Args(name))
    torch.cuda.synchronize()
    _same(got, want, f"{name} {shape}")
    # the caller-owned-buffer form writes the same bytes and allocates nothing
    bufs = ops.compress_outputs(w.shape, Args(name), w.dtype, w.device)
    got2 = ops.compress_weight(w, Args(name), out=bufs)
    for k in ("weight_packed", "weight", "weight_scale", "weight_zero_point"):
        if k in got:
            assert got2[k].data_ptr() == bufs[k].data_ptr(), f"{k} was not written into the caller's buffer"
            assert_bits_equal(got2[k], got[k], f"{name} out=:{k}")


def test_fullsize_int4_stack_bit_exact_vs_ct_cuda():
    """A stacked launch (the bench's arena form: [units, rows, cols], one launch) against per-matrix CT, 4 Qwen3-4B gate/up matrices
    + MiniMax expert shapes."""
    from quantizers_b200 import ops

    for name, (E, R, C) in (("int4_g128_asym", (4, 9728, 2560)), ("int4_g32_sym", (6, 1536, 3072)), ("int4_g32_sym", (6, 3072, 1536))):
        fmt, a = L.format_args(name)
        w = torch.stack([_w(R, C, 77 + e) for e in range(E)])
        got = ops.compress_weight(w, Args(name))
        for e in range(E):
            want = L.compress(w[e], fmt, a)
            _same({k: (v[e] if k != "weight_shape" else v) for k, v in got.items()}, want, f"{name} stack[{e}]")


@pytest.mark.parametrize("experts", [128])
def test_nvfp4_expert_stacks_bit_exact_vs_ct_cuda(experts):
    """config 4: one Qwen3-30B-A3B layer, all 128 experts.  gate/up of an expert share min(global_scale)
    (LLMC update_fused_layer_weight_global_scales); CT runs per matrix with that fused scale."""
    from quantizers_b200 import ops

    fmt, a = L.format_args("nvfp4")
    args = Args("nvfp4")
    gen = torch.Generator(device="cuda").manual_seed(99)
    # per-expert magnitude spread so that the global scales differ between siblings and experts
    mag = 0.02 * (0.25 + 4 * torch.rand(experts * 2, 1, 1, generator=gen, device="cuda"))
    gate_up = (torch.randn(experts * 2, 768, 2048, generator=gen, device="cuda") * mag).to(torch.bfloat16)
    gate_up[3, 5, :32] = 0.0
    gate_up[4, 7, :] = -0.0
    down = (torch.randn(experts, 2048, 768, generator=gen, device="cuda") * 0.02).to(torch.bfloat16)
    got_gu = ops.compress_weight(gate_up, args, fuse_span=2)
    got_d = ops.compress_weight(down, args)
    torch.cuda.synchronize()
    for e in range(experts):
        gs = torch.minimum(L.global_scale(gate_up[2 * e]), L.global_scale(gate_up[2 * e + 1]))
        for j in range(2):
            want = L.compress(gate_up[2 * e + j], fmt, a, gs=gs)
            _same({k: v[2 * e + j] for k, v in got_gu.items()}, want, f"nvfp4 gate/up expert {e}[{j}]")
        want = L.compress(down[e], fmt, a)
        _same({k: v[e] for k, v in got_d.items()}, want, f"nvfp4 down expert {e}")


def test_nvfp4_dense_shapes_bit_exact_vs_ct_cuda():
    """recipe_Dense_NVFP4.yaml on Qwen3-4B MLP shapes (q/k/v- and gate/up-fused global scales)."""
    from quantizers_b200 import ops

    fmt, a = L.format_args("nvfp4")
    args = Args("nvfp4")
    gu = torch.stack([_w(9728, 2560, 5), _w(9728, 2560, 6) * 2])
    got = ops.compress_weight(gu, args, fuse_span=2)
    gs = torch.minimum(L.global_scale(gu[0]), L.global_scale(gu[1]))
    for j in range(2):
        _same({k: v[j] for k, v in got.items()}, L.compress(gu[j], fmt, a, gs=gs), f"nvfp4 dense gate/up[{j}]")
    d = _w(2560, 9728, 7)
    _same(ops.compress_weight(d, args), L.compress(d, fmt, a), "nvfp4 dense down")


def test_arena_pass_equals_ct_on_a_layer():
    """One Qwen3-4B decoder layer through the scheduler's arena path with caller-owned outputs (what bench.py times) == CT per
    matrix."""
    from quantizers_b200 import scheduler as S

    spec = S.qwen3_4b(layers=1)
    arena = S.build_arena(spec, [3], "cuda")
    bufs = S.alloc_outputs(spec, arena)
    res = S.quantize_arena(spec, arena, out=bufs)
    res = S.quantize_arena(spec, arena, out=bufs)  # second pass into the same buffers
    torch.cuda.synchronize()
    names = {"FP8_BLOCK": "fp8_block", "W4A16_ASYM": "int4_g128_asym"}
    for m in spec.matrices:
        fmt, a = L.format_args(names[m.preset])
        for j in range(m.per_unit):
            want = L.compress(arena[m.name][j], fmt, a)
            _same({k: (v[j] if k != "weight_shape" else v) for k, v in res[m.name].items()}, want, f"arena {m.name}[{j}]")


@pytest.mark.parametrize("model", ["qwen3_4b", "qwen3_30b_a3b", "glm47_flash"])
def test_concurrent_class_launches_and_graph_replay_equal_the_sequential_pass(model):
    """What the strong-scaling bench legs time: the classes of a shard launched on side streams (fork / join), captured once in a CUDA
    graph and replayed -- must leave the same bytes as the plain sequential pass, for the (class, unit)-balanced partition of rank 1 of 3."""
    from quantizers_b200 import scheduler as S

    spec = {"qwen3_4b": S.qwen3_4b(layers=1), "qwen3_30b_a3b": S.qwen3_30b_a3b(layers=1, experts=1), "glm47_flash": S.glm47_flash(units=1)}[model]
    mine = S.partition_balanced(spec, 5, 3)[1]
    arena = S.build_arena_classes(spec, mine, "cuda")
    want = S.quantize_arena(spec, arena)
    bufs = S.alloc_outputs(spec, arena)
    S.quantize_arena(spec, arena, out=bufs, concurrent=True)      # warm-up outside the capture
    torch.cuda.synchronize()
    for b in bufs.values():
        for k, v in b.items():
            if torch.is_tensor(v) and v.is_cuda and not k.startswith("_"):
                v.view(torch.uint8).zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        S.quantize_arena(spec, arena, out=bufs, concurrent=True)
    graph.replay()
    torch.cuda.synchronize()
    assert set(want) == set(mine)
    for name, w in want.items():
        for k, v in w.items():
            if torch.is_tensor(v) and v.is_cuda:
                assert torch.equal(bufs[name][k].view(torch.uint8), v.view(torch.uint8)), (name, k)
