"""GPU: throughput of the entry points that take CALLER-SUPPLIED qparams (SURVEY.md §8a Q3 quantize, Q4 fake_quantize, Q5 dequantize and
the compressors' quantize_pack) -- the path the registered compressors take when an observer has already written weight_scale /
weight_zero_point.  One [8 x 9728, 2560] bf16 matrix (398 MB).  Algorithmic bytes: input + output, qparams negligible.
Writes gpurun_out/qparams_paths.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
peak = 6549.4
w = synth_stack(list(range(8)), 9728, 2560, 0, dev).reshape(-1, 2560).contiguous()
n = w.numel()
names = [a for a in sys.argv[1:] if a != "PACK"] or ([] if sys.argv[1:] else ["W4A16", "W4A16_ASYM", "INT4_G32_SYM", "FP8_BLOCK", "FP8_CHANNEL", "FP8_G32", "NVFP4"])
rows = []


def timed(label, nbytes, fn):
    try:
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"{label:40s} ERROR {str(e)[:160]}", flush=True)
        return None
    ms = e0.elapsed_time(e1) / 10
    gbps = nbytes / ms / 1e6
    rows.append({"op": label, "ms": ms, "algorithmic_GBps": gbps, "frac_of_hbm_peak": gbps / peak})
    print(f"{label:40s} {ms*1e3:8.1f} us  {gbps:6.0f} GB/s algorithmic  {gbps/peak:.3f}", flush=True)
    return out


for name in names:
    a = PRESETS[name]
    gsc = ops.weight_global_scales(w) if a.num_bits == 4 and a.type == "float" else None
    mn, mx = (None, None)
    if gsc is None:
        mn, mx = ops.observe_minmax(w, a)
        scale, zp = ops.calculate_qparams(mn, mx, a)
    else:
        sd = ops.compress_weight(w, a)
        scale, zp = sd["weight_scale"], None
        gsc = sd["weight_global_scale"]
    if a.symmetric:
        zp = None
    code_bytes = n * (0.5 if a.num_bits == 4 else 1.0)
    # Q3 quantize: bf16 -> codes (INT: int8 per element; FP8: e4m3; FP4: grid values in bf16)
    q_out_bytes = n * (2 if (a.type == "float" and a.num_bits == 4) else 1)
    q = timed(f"{name} quantize", n * 2 + q_out_bytes,
              lambda: ops.quantize(w, scale, zp, a, dtype=(torch.int8 if a.type == "int" else (torch.float8_e4m3fn if a.num_bits == 8 else None)), global_scale=gsc))
    timed(f"{name} fake_quantize", n * 4, lambda: ops.fake_quantize(w, scale, zp, a, global_scale=gsc))
    if a.strategy in ("group", "tensor_group"):
        timed(f"{name} quantize_pack", n * 2 + code_bytes, lambda: ops.quantize_pack(w, scale, zp, a, global_scale=gsc))
    if q is not None and not (a.type == "float" and a.num_bits == 4):
        timed(f"{name} dequantize", q.numel() * q.element_size() + n * 2,
              lambda: ops.dequantize(q, scale, zp, args=a, dtype=torch.bfloat16))
if not sys.argv[1:] or "PACK" in sys.argv[1:]:
    # stand-alone pack helpers (Q7 / Q8 / Q9): only reached when a caller packs codes it already holds -- the patched compressors
    # go through quantize_pack / the fused decompress kernels instead
    codes = torch.randint(-8, 8, w.shape, dtype=torch.int8, device=dev)
    packed = timed("pack_to_int32 (int8 codes, 4 bit)", n * 1.5, lambda: ops.pack_to_int32(codes, 4))
    timed("unpack_from_int32 (4 bit)", n * 1.5, lambda: ops.unpack_from_int32(packed, 4, w.shape))
    vals = ops.quantize(w[:, :], *(lambda sd: (sd["weight_scale"].to(torch.bfloat16), None))(ops.compress_weight(w, PRESETS["NVFP4"])), PRESETS["NVFP4"],
                        global_scale=ops.weight_global_scales(w))
    p4 = timed("pack_fp4_to_uint8 (bf16 grid values)", n * 2.5, lambda: ops.pack_fp4_to_uint8(vals))
    timed("unpack_fp4_from_uint8 (-> bf16)", n * 2.5, lambda: ops.unpack_fp4_from_uint8(p4, w.shape[0], w.shape[1], torch.bfloat16))
os.makedirs("gpurun_out", exist_ok=True)
tag = os.environ.get("B200Q_BENCH_TAG", "")
json.dump({"matrix": list(w.shape), "rows": rows}, open(f"gpurun_out/qparams_paths{'_' + tag if tag else ''}.json", "w"), indent=1)
