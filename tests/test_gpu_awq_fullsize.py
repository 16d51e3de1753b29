"""GPU: BASELINE.json config 1 at full size -- the AWQ scale search of one whole Qwen3-4B decoder layer (T = 64 x 512 calibration
tokens, n_grid 20, duo_scaling, default Llama/Qwen mappings) against the restated llmcompressor loop evaluated with LIVE
compressed-tensors arithmetic on this GPU (oracle/llmc_live.py), sample by sample like ``_run_samples``.

Per mapping (q/k/v with the self_attn parent, gate/up with the mlp parent, down_proj with itself as parent):
  * the 20 per-ratio losses agree within 1e-3 relative (north_star's tolerance; the fused tcgen05 kernels never materialise outputs),
  * the argmin ratio is the same (first minimum wins, like upstream's ``if loss < best_error`` scan),
  * the best scale vectors agree to fp32 round-off,
and after smoothing with those scales the final RTN ``pack-quantized`` tensors are bit-identical to ``Compressor.compress`` of live
CT on the same smoothed weights (W5).  Anchors: REF:configs/test-quantize_qwen3-4b-awq.yaml, REF:configs/recipes/recipe_awq_w4a16.yaml:13-32,
SURVEY.md Appendix A 488-509.
"""
import json
import os

import pytest
import torch

from oracle import ct_live as L
from tests.test_gpu_compress import Args
from tests.util import assert_bits_equal

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not L.available(), reason="compressed_tensors is not importable")]

CFG = dict(n_heads=32, n_kv=8, head_dim=128, seq_len=512)


def _layer(T):
    from quantizers_b200 import scheduler as S

    return S.synth_awq_layer(1, T, torch.device("cuda", 0))


def _mappings(w, acts):
    from oracle import llmc_live as V
    from quantizers_b200 import awq

    return [
        ("qkv", "attn_in", ["q", "k", "v"], "input_layernorm",
         lambda: awq.AttentionParent(w["o"], CFG["n_heads"], CFG["n_kv"], CFG["head_dim"], CFG["seq_len"], w["q_norm"], w["k_norm"]),
         lambda: V.attention_parent(w["o"], CFG["n_heads"], CFG["n_kv"], CFG["head_dim"], w["q_norm"], w["k_norm"])),
        ("gate_up", "mlp_in", ["gate", "up"], "post_attention_layernorm", lambda: awq.MLPParent(w["down"]), lambda: V.mlp_parent(w["down"])),
        ("down", "down_in", ["down"], "up", lambda: awq.linear_parent, lambda: V.linear_parent),
    ]


@pytest.mark.parametrize("scheme", ["int4_g128_asym", "int4_g32_sym"])
def test_config1_layer_search_matches_live_ct(scheme):
    from oracle import llmc_live as V
    from quantizers_b200 import awq, ops

    T = 64 * 512
    w, acts = _layer(T)
    fmt, ct_args = L.format_args(scheme)
    args = Args(scheme)
    S = CFG["seq_len"]
    report = {}
    for name, act_key, balance, smooth_key, ours_parent, ref_parent in _mappings(w, acts):
        x = acts[act_key]
        ws = [w[k] for k in balance]
        s, r, losses = awq.compute_best_scale(x, ws, ours_parent(), args)
        batches = [x[t0:t0 + S] for t0 in range(0, T, S)]
        s_ref, r_ref, l_ref = V.compute_best_scale(batches, ws, ref_parent(), ct_args)
        rel = max(abs(a - b) / b for a, b in zip(losses, l_ref))
        report[name] = {"ratio": r, "ratio_ref": r_ref, "max_rel_loss_diff": rel}
        assert rel < 1e-3, f"{scheme} {name}: per-ratio losses differ by {rel:.2e} relative"
        assert r == r_ref, f"{scheme} {name}: argmin ratio {r} != reference {r_ref} (losses {losses} vs {l_ref})"
        assert torch.allclose(s, s_ref.cpu().float(), rtol=2e-5, atol=0), f"{scheme} {name}: best scales differ"
        # _smooth with OUR scales on both copies (bit-compared), then the next mapping sees identical weights
        ref_copy = [t.clone() for t in ws]
        ref_smooth = w[smooth_key].clone()
        V.smooth(ref_copy, ref_smooth, s.to(x.device))
        awq.smooth(ws, w[smooth_key], s)
        for a, b, k in zip(ws, ref_copy, balance):
            assert_bits_equal(a, b, f"{scheme} {name}: smoothed {k}")
        assert_bits_equal(w[smooth_key], ref_smooth, f"{scheme} {name}: smoothed {smooth_key}")
    # W5: final RTN of the smoothed weights, fused compress vs live CT, bit-exact
    for k in ("q", "k", "v", "gate", "up", "down"):
        got = ops.compress_weight(w[k], args)
        want = L.compress(w[k], fmt, ct_args)
        for key, v in want.items():
            if key == "weight_shape":
                continue
            assert_bits_equal(got[key], v, f"{scheme} final {k}:{key}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open(f"gpurun_out/awq_fullsize_{scheme}.json", "w"), indent=1)


def test_fidelity_switches_report():
    """The two places where llmcompressor versions are known to differ (x_mean accumulation dtype, loss form) are one flag away in
    both the product and the oracle; this reports how often the argmin moves between the conventions on config-1 shapes
    (8 192 tokens to bound the run time of the un-fused 'mse_bf16' path) and checks product == oracle under every convention."""
    from oracle import llmc_live as V
    from quantizers_b200 import awq

    T = 16 * 512
    w, acts = _layer(T)
    fmt, ct_args = L.format_args("int4_g128_asym")
    args = Args("int4_g128_asym")
    S = CFG["seq_len"]
    table = {}
    for name, act_key, balance, smooth_key, ours_parent, ref_parent in _mappings(w, acts):
        x = acts[act_key]
        ws = [w[k] for k in balance]
        batches = [x[t0:t0 + S] for t0 in range(0, T, S)]
        for xm in ("fp32", "act"):
            for lf in ("float_pow", "mse_bf16"):
                s, r, losses = awq.compute_best_scale(x, ws, ours_parent(), args, x_mean_dtype=xm, loss_form=lf, sample_len=S, token_chunk=S)
                s_ref, r_ref, l_ref = V.compute_best_scale(batches, ws, ref_parent(), ct_args, x_mean_dtype=xm, loss_form=lf)
                rel = max(abs(a - b) / b for a, b in zip(losses, l_ref))
                table[f"{name}:{xm}:{lf}"] = {"ratio": r, "ratio_ref": r_ref, "max_rel_loss_diff": rel}
                # bf16-rounded per-batch sums quantise the loss to ~2^-9 relative: the tolerance of the bf16 forms is the format's
                tol = 1e-3 if lf == "float_pow" else 8e-3
                assert rel < tol, f"{name} {xm} {lf}: losses differ by {rel:.2e}"
    base = {m: table[f"{m}:fp32:float_pow"]["ratio"] for m in ("qkv", "gate_up", "down")}
    moved = {k: v["ratio"] for k, v in table.items() if v["ratio"] != base[k.split(":")[0]]}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"table": table, "default_argmin": base, "argmin_moved_under": moved}, open("gpurun_out/awq_fidelity_switches.json", "w"), indent=1)
