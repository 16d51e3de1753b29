"""GPU: decode throughput (compressed tensors -> bf16 weights) per format on one 32 x [9728, 2560] stack; gpurun_out/decompress.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, SchemeArgs, synth_stack

dev = torch.device("cuda", 0)
peak = 6549.4
E, R, C = (int(v) for v in os.environ.get("B200Q_BENCH_SHAPE", "32,9728,2560").split(","))
w = synth_stack(list(range(E)), R, C, 0, dev)
rows = []
names = sys.argv[1:] or ["W4A16", "W4A16_ASYM", "INT4_G32_SYM", "INT4_G32_ASYM", "NVFP4", "FP8_BLOCK", "FP8_G32", "FP8_CHANNEL"]
for name in names:
    a = PRESETS[name] if name in PRESETS else SchemeArgs(4, "int", False, "group", 32)  # INT4_G32_ASYM
    sd = ops.compress_weight(w, a)
    if a.type == "int":
        fn = lambda: ops.decompress_int_packed(sd["weight_packed"], sd["weight_scale"], sd.get("weight_zero_point"), w.shape, a)
    elif a.num_bits == 4:
        fn = lambda: ops.decompress_nvfp4(sd["weight_packed"], sd["weight_scale"], sd["weight_global_scale"])
    else:
        # dequantize() is 2-D like the reference's; rows of a stack are independent for these strategies, so one call on the stacked
        # rows measures the kernel rather than 32 Python round trips
        codes2d, scale2d = sd["weight"].reshape(-1, w.shape[-1]), sd["weight_scale"].reshape(-1, sd["weight_scale"].shape[-1])
        fn = lambda: ops.dequantize(codes2d, scale2d, args=a, dtype=torch.bfloat16)
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
        print(name, "ERROR", str(e)[:200], flush=True)
        continue
    ms = e0.elapsed_time(e1) / 10
    in_bytes = sum(v.numel() * v.element_size() for k, v in sd.items() if k != "weight_shape")
    alg = (in_bytes + w.numel() * 2) / ms / 1e6
    rows.append({"scheme": name, "ms": ms, "bf16_out_GBps": w.numel() * 2 / ms / 1e6, "algorithmic_GBps": alg, "frac_of_hbm_peak": alg / peak})
    print(f"{name:13s} {ms*1e3:8.1f} us  {w.numel()*2/ms/1e6:6.0f} GB/s bf16-out  {alg:6.0f} GB/s algorithmic  {alg/peak:.3f}", flush=True)
    del out, sd
os.makedirs("gpurun_out", exist_ok=True)
tag = os.environ.get("B200Q_BENCH_TAG") or os.environ.get("B200Q_DECODE_INT4_PRE", "")  # second name: the tag the round-1 sweep scripts used
json.dump({"stack": list(w.shape), "rows": rows}, open(f"gpurun_out/decompress{'_' + tag if tag else ''}.json", "w"), indent=1)
