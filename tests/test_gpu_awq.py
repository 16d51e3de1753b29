"""GPU: AWQ statistics kernels and the scale search against the restated oracle (loss curve within 1e-3 relative,
same argmin ratio -- BASELINE.json north_star tolerance), plus activation observers."""
import pytest
import torch

from oracle import llmc_restated as R
from oracle import oracle as O
from tests.test_gpu_compress import Args
from tests.util import assert_bits_equal

pytestmark = pytest.mark.gpu


def _problem(T=1024, K=512, N=256, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(T, K, generator=g) * (1 + 3 * torch.rand(K, generator=g))).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.02).to(torch.bfloat16)
    w[:, 7] *= 20
    return x, w


def test_abs_sum_and_wmean():
    from quantizers_b200 import awq

    x, w = _problem()
    want = O.abs_sum_cols(x)
    got = awq.abs_sum_cols(x.cuda()).double().cpu()
    assert torch.allclose(got, want, rtol=1e-5)
    for g in (32, 128):
        wm = awq.compute_layer_means([w.cuda(), w.flip(0).cuda()], g).cpu()
        ref = O.w_mean([w, w.flip(0)], g).float()
        assert torch.equal(wm, ref), g  # fp64 accumulation of bf16 values is exact => order independent
    # empty activation batch (unrouted expert) is skipped
    acc = awq.abs_sum_cols(torch.zeros(0, 512, dtype=torch.bfloat16).cuda())
    assert float(acc.sum()) == 0.0


def test_awq_scales_and_scaled_fake_quantize():
    from quantizers_b200 import awq

    x, w = _problem(seed=2)
    xm, _ = R.accumulate_abs_mean([x])
    for name, geom, qtype, nb, sym in (("int4_g32_sym", O.Geom(O.GROUP, 32), O.INT, 4, True),
                                       ("int4_g128_asym", O.Geom(O.GROUP, 128), O.INT, 4, False),
                                       ("fp8_g32", O.Geom(O.GROUP, 32), O.FP8, 8, True)):
        wm = R.compute_layer_means([w], geom.group)
        ratios = [i / 20 for i in range(20)]
        got = awq.awq_scales(xm.cuda(), wm.cuda(), ratios, True).cpu()
        for i, r in enumerate(ratios):
            ref = R.awq_scales(xm, wm, r, True)
            assert torch.allclose(got[i], ref, rtol=2e-6, atol=0), (name, r)  # powf vs torch.pow: ulp-level
        # identical scales in -> bit-identical weights out
        s = R.awq_scales(xm, wm, 0.35, True)
        out = awq.scaled_fake_quantize(w.cuda(), s.cuda(), Args(name))
        assert_bits_equal(out, R.scaled_fake_quantize(w, s, geom, qtype, nb, sym), name)


@pytest.mark.parametrize("name,geom,qtype,nb,sym", [("int4_g32_sym", O.Geom(O.GROUP, 32), O.INT, 4, True),
                                                    ("int4_g128_asym", O.Geom(O.GROUP, 128), O.INT, 4, False)])
def test_compute_best_scale_linear_parent(name, geom, qtype, nb, sym):
    from quantizers_b200 import awq

    x, w = _problem(T=768, K=512, N=256, seed=4)
    s_ref, r_ref, l_ref = R.compute_best_scale([x[:384], x[384:]], [w], R.linear_parent, geom, qtype, nb, sym)
    s, r, l = awq.compute_best_scale(x.cuda(), [w.cuda()], awq.linear_parent, Args(name), token_chunk=500)
    rel = [abs(a - b) / b for a, b in zip(l, l_ref)]
    assert max(rel) < 1e-3, rel          # per-ratio losses within 1e-3 relative
    assert r == r_ref                     # same argmin ratio (first minimum wins)
    assert torch.allclose(s, s_ref, rtol=1e-5)


def test_compute_best_scale_mlp_parent():
    from quantizers_b200 import awq

    g = torch.Generator().manual_seed(9)
    x, gate = _problem(T=512, K=256, N=384, seed=6)
    up = (torch.randn(384, 256, generator=g) * 0.02).to(torch.bfloat16)
    down = (torch.randn(256, 384, generator=g) * 0.02).to(torch.bfloat16)
    geom = O.Geom(O.GROUP, 32)
    s_ref, r_ref, l_ref = R.compute_best_scale([x], [gate, up], R.mlp_parent(down), geom, O.INT, 4, True)
    s, r, l = awq.compute_best_scale(x.cuda(), [gate.cuda(), up.cuda()], awq.mlp_parent(down.cuda()), Args("int4_g32_sym"))
    rel = [abs(a - b) / b for a, b in zip(l, l_ref)]
    assert max(rel) < 1e-3, rel
    assert r == r_ref


def test_activation_global_scale_running():
    """static_minmax input_global_scale: running min/max over batches == one pass over the concatenation."""
    from quantizers_b200 import ops

    x, _ = _problem(seed=8)
    xd = x.cuda()
    state = torch.tensor([float("inf"), float("-inf")], device="cuda")
    for lo, hi in ((0, 100), (100, 100), (100, 1024)):
        if hi > lo:
            gs = ops.observe_global_scale(xd[lo:hi], state)
    want = R.activation_global_scale([x[:100], x[100:]])
    assert gs.item() == want.item()
    assert ops.observe_global_scale(xd).item() == want.item()
    mn, mx = float(x.float().min()), float(x.float().max())
    assert state[0].item() == mn and state[1].item() == mx


def _torch_gemm_loss(x, w_ref, w_q):
    """Plain torch restatement of _run_samples + _compute_loss for a single-Linear parent (bf16 GEMM, fp32 accumulate,
    bf16 outputs, bf16 difference squared and summed in fp32/fp64)."""
    ref = torch.nn.functional.linear(x, w_ref)
    out = []
    for r in range(w_q.shape[0]):
        d = ref - torch.nn.functional.linear(x, w_q[r])
        out.append(d.float().pow(2).double().sum())
    return torch.stack(out)


@pytest.mark.parametrize("T,K,N,R", [(256, 128, 256, 2), (1024, 512, 256, 3), (1000, 520, 300, 2), (4096, 2560, 1024, 4),
                                     (77, 64, 40, 1), (20000, 256, 512, 2)])
def test_gemm_loss_fused_vs_torch(T, K, N, R):
    """tcgen05 fused loss GEMM against torch bf16 matmul on the same device; ragged M/N/K tiles included."""
    from quantizers_b200 import awq

    g = torch.Generator().manual_seed(T + K + N)
    x = (torch.randn(T, K, generator=g) * (1 + 3 * torch.rand(K, generator=g))).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) * 0.02).to(torch.bfloat16).cuda()
    wq = torch.stack([(w.float() + 0.002 * (r + 1) * torch.randn(N, K, generator=g).cuda()).to(torch.bfloat16) for r in range(R)])
    got = awq.gemm_loss_fused(x, w, wq).double().cpu()
    want = _torch_gemm_loss(x, w, wq).cpu()
    rel = ((got - want).abs() / want).max().item()
    assert rel < 1e-3, (got, want)
    # identical weights -> exactly zero loss (same accumulation order for both operands)
    z = awq.gemm_loss_fused(x, w, w[None].contiguous())
    assert float(z[0]) == 0.0


@pytest.mark.parametrize("name,geom,qtype,nb,sym", [("int4_g32_sym", O.Geom(O.GROUP, 32), O.INT, 4, True),
                                                    ("int4_g128_asym", O.Geom(O.GROUP, 128), O.INT, 4, False)])
def test_compute_best_scale_fused_linear(name, geom, qtype, nb, sym):
    """Whole search through the fused tensor-core loss kernel against the restated CPU oracle."""
    from quantizers_b200 import awq

    x, w = _problem(T=768, K=512, N=256, seed=4)
    s_ref, r_ref, l_ref = R.compute_best_scale([x[:384], x[384:]], [w], R.linear_parent, geom, qtype, nb, sym)
    s, r, l = awq.compute_best_scale(x.cuda(), [w.cuda()], awq.linear_parent, Args(name), fused_linear=True)
    rel = [abs(a - b) / b for a, b in zip(l, l_ref)]
    assert max(rel) < 1e-3, rel
    assert r == r_ref
    assert torch.allclose(s, s_ref, rtol=1e-5)


@pytest.mark.parametrize("T,K,N,V,swiglu", [(256, 128, 256, 2, False), (1000, 520, 304, 3, False), (4096, 2560, 1024, 2, False),
                                            (256, 128, 128, 2, True), (1000, 520, 200, 3, True), (4096, 2560, 1280, 2, True)])
def test_gemm_project_vs_torch(T, K, N, V, swiglu):
    """tcgen05 projection (plain and SwiGLU epilogue) against torch bf16 ops; ragged M/N/K tiles included."""
    from quantizers_b200 import awq

    F = torch.nn.functional
    g = torch.Generator().manual_seed(T + K + N)
    x = torch.randn(T, K, generator=g).to(torch.bfloat16).cuda()
    rows = 2 * N if swiglu else N
    w = (torch.randn(V, rows, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    got = awq.gemm_project(x, w, swiglu=swiglu)
    assert got.shape == (V, T, N)
    for v in range(V):
        want = F.silu(F.linear(x, w[v, :N])) * F.linear(x, w[v, N:]) if swiglu else F.linear(x, w[v])
        diff = (got[v].float() - want.float()).abs()
        # fp32 accumulation order differs from cuBLAS: allow one bf16 ulp on a small fraction of the outputs
        tol = want.float().abs() * 2 ** -7 + 1e-6
        assert bool((diff <= tol).all()), (v, float(diff.max()))
        assert float((diff > 0).float().mean()) < 0.05


def test_gemm_loss_pairs_varying_a():
    from quantizers_b200 import awq

    g = torch.Generator().manual_seed(11)
    T, K, N, R = 1000, 520, 300, 3
    a = torch.randn(T, K, generator=g).to(torch.bfloat16).cuda()
    aq = torch.stack([(a.float() + 0.01 * (r + 1) * torch.randn(T, K, generator=g).cuda()).to(torch.bfloat16) for r in range(R)])
    b = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    got = awq.gemm_loss_pairs(a, aq, b, None).double().cpu()
    ref = torch.nn.functional.linear(a, b)
    want = torch.stack([(ref - torch.nn.functional.linear(aq[r], b)).float().pow(2).double().sum() for r in range(R)]).cpu()
    assert ((got - want).abs() / want).max().item() < 1e-3


def test_compute_best_scale_fused_mlp_parent():
    from quantizers_b200 import awq

    g = torch.Generator().manual_seed(9)
    x, gate = _problem(T=512, K=256, N=384, seed=6)
    up = (torch.randn(384, 256, generator=g) * 0.02).to(torch.bfloat16)
    down = (torch.randn(256, 384, generator=g) * 0.02).to(torch.bfloat16)
    geom = O.Geom(O.GROUP, 32)
    s_ref, r_ref, l_ref = R.compute_best_scale([x], [gate, up], R.mlp_parent(down), geom, O.INT, 4, True)
    s, r, l = awq.compute_best_scale(x.cuda(), [gate.cuda(), up.cuda()], awq.MLPParent(down.cuda()), Args("int4_g32_sym"))
    rel = [abs(a - b) / b for a, b in zip(l, l_ref)]
    assert max(rel) < 1e-3, rel
    assert r == r_ref


def test_compute_best_scale_fused_attention_parent():
    """input_layernorm -> q/k/v mapping with the self-attention parent (Qwen3: q/k RMSNorm, RoPE, causal GQA, o_proj)."""
    from quantizers_b200 import awq

    g = torch.Generator().manual_seed(21)
    H, HKV, D, S, B, K = 4, 2, 64, 128, 4, 256
    x = (torch.randn(B * S, K, generator=g) * (1 + 3 * torch.rand(K, generator=g))).to(torch.bfloat16)
    mk = lambda n, k, s=0.03: (torch.randn(n, k, generator=g) * s).to(torch.bfloat16)
    wq_, wk_, wv_, wo_ = mk(H * D, K), mk(HKV * D, K), mk(HKV * D, K), mk(K, H * D)
    qn, kn = (1 + 0.1 * torch.randn(D, generator=g)).to(torch.bfloat16), (1 + 0.1 * torch.randn(D, generator=g)).to(torch.bfloat16)
    geom = O.Geom(O.GROUP, 128)
    par = R.attention_parent(wo_, H, HKV, D, qn, kn)
    s_ref, r_ref, l_ref = R.compute_best_scale([x[i * S:(i + 1) * S] for i in range(B)], [wq_, wk_, wv_], par, geom, O.INT, 4, False)
    gp = awq.AttentionParent(wo_.cuda(), H, HKV, D, S, qn.cuda(), kn.cuda())
    res = {}
    for fused in (True, False):
        res[fused] = awq.compute_best_scale(x.cuda(), [wq_.cuda(), wk_.cuda(), wv_.cuda()], gp, Args("int4_g128_asym"), fused=fused)
    # tensor-core evaluation vs the same parent through torch GEMMs on the same device: the north-star tolerance
    rel = [abs(a - b) / b for a, b in zip(res[True][2], res[False][2])]
    assert max(rel) < 1e-3, rel
    assert res[True][1] == res[False][1]
    # vs the CPU oracle: the softmax(QK^T)V core is a different implementation (flash attention in bf16 on the GPU, explicit
    # fp32 softmax in the oracle), which alone moves the per-ratio loss by ~2e-3 at this size; the argmin must still agree
    for fused in (True, False):
        rel = [abs(a - b) / b for a, b in zip(res[fused][2], l_ref)]
        assert max(rel) < 5e-3, (fused, rel)
        assert res[fused][1] == r_ref


@pytest.mark.parametrize("H,HKV,D,S,B", [(4, 2, 64, 64, 3), (32, 8, 128, 512, 2), (5, 1, 128, 70, 1), (3, 1, 64, 33, 5)])
def test_qk_norm_rope_vs_eager(H, HKV, D, S, B):
    """In-place q/k RMSNorm + RoPE kernel against the eager bf16 op chain of transformers' Qwen3Attention."""
    from quantizers_b200 import awq

    g = torch.Generator().manual_seed(5)
    T = B * S
    qkv = torch.randn(T, (H + 2 * HKV) * D, generator=g).to(torch.bfloat16).cuda()
    qn = (1 + 0.1 * torch.randn(D, generator=g)).to(torch.bfloat16).cuda()
    kn = (1 + 0.1 * torch.randn(D, generator=g)).to(torch.bfloat16).cuda()
    par = awq.AttentionParent(torch.zeros(8, H * D, dtype=torch.bfloat16).cuda(), H, HKV, D, S, qn, kn)

    def rms(x, w):
        v = x.float()
        v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-6)
        return w * v.to(x.dtype)

    def rope(x):  # x [B, S, h, D]
        cos, sin = par.cos[None, :, None, :], par.sin[None, :, None, :]
        x1, x2 = x[..., : D // 2], x[..., D // 2:]
        return x * cos + torch.cat((-x2, x1), dim=-1) * sin

    q, k, v = qkv.split([H * D, HKV * D, HKV * D], dim=-1)
    q_ref = rope(rms(q.reshape(B, S, H, D), qn)).reshape(T, H * D)
    k_ref = rope(rms(k.reshape(B, S, HKV, D), kn)).reshape(T, HKV * D)
    want = torch.cat([q_ref, k_ref, v], dim=-1)
    got = qkv.clone()
    from quantizers_b200 import _lib as L
    import ctypes
    L.check(L.lib().b200q_qk_norm_rope(L.ptr(got), T, H, HKV, D, S, L.ptr(qn), L.ptr(kn), L.ptr(par.cos), L.ptr(par.sin), ctypes.c_float(1e-6),
                                       L.stream_ptr(got.device)))
    assert torch.equal(got[:, (H + HKV) * D:], v)  # v columns untouched
    diff = (got.float() - want.float()).abs()
    # the only freedom is the summation order of mean(x^2): it flips the bf16 rounding of x * rsqrt(..) on a tiny fraction of the
    # elements, and that one ulp propagates through w * vn and y * cos + rot * sin (a CPU run of the eager chain with an fp64 mean
    # instead of the fp32 one shows the same: 3e-6 of the elements move, by up to 2 ulps)
    assert bool((diff <= want.float().abs() * 2 ** -6 + 2 ** -6).all()), float(diff.max())
    assert float((diff > 0).float().mean()) < 0.002


def test_search_expert_mappings_matches_oracle():
    """config 5 (ii): per-expert w3 -> w2 mappings of a MoE layer, each an independent single-Linear search; the best scales
    are folded in like _smooth (w2 *= s, w3 rows /= s).  Losses within 1e-3 relative, same argmin, per expert."""
    from quantizers_b200 import awq

    E, T, H, I = 3, 512, 256, 384   # experts, tokens, hidden, expert intermediate
    g = torch.Generator().manual_seed(17)
    x = (torch.randn(T, H, generator=g) * (1 + 3 * torch.rand(H, generator=g))).to(torch.bfloat16)
    w1 = (torch.randn(E, I, H, generator=g) * 0.05).to(torch.bfloat16)
    w3 = (torch.randn(E, I, H, generator=g) * 0.05).to(torch.bfloat16)
    w2 = (torch.randn(E, H, I, generator=g) * 0.02).to(torch.bfloat16)
    F = torch.nn.functional
    xs = [F.silu(F.linear(x, w1[e])) * F.linear(x, w3[e]) for e in range(E)]
    geom = O.Geom(O.GROUP, 32)
    w2_dev, w3_dev = w2.cuda(), w3.cuda()
    got = awq.search_expert_mappings([v.cuda() for v in xs], w2_dev, Args("int4_g32_sym"), smooth_weight=w3_dev)
    assert len(got) == E
    for e in range(E):
        s_ref, r_ref, l_ref = R.compute_best_scale([xs[e]], [w2[e]], R.linear_parent, geom, O.INT, 4, True)
        s, r, l = got[e]
        assert max(abs(a - b) / b for a, b in zip(l, l_ref)) < 1e-3, e
        assert r == r_ref, e
        assert torch.allclose(s, s_ref, rtol=1e-5)
        new_w, new_s = R.smooth([w2[e]], w3[e], s)  # fold the SAME scales in on the CPU
        assert_bits_equal(w2_dev[e], new_w[0], f"w2[{e}] smoothed")
        assert_bits_equal(w3_dev[e], new_s, f"w3[{e}] smoothed")
    with pytest.raises(ValueError):
        awq.search_expert_mappings([xs[0].cuda()], w2_dev, Args("int4_g32_sym"))


def test_search_moe_block_mapping_matches_oracle():
    """config 5 (i): ONE scale vector for all experts' w1 / w3, parent = the routed sparse-MoE block, loss on its output.
    Routing (top-2 of 6 experts, one expert never routed) is computed once and handed to both sides."""
    from quantizers_b200 import awq

    E, T, H, I, K = 6, 640, 256, 384, 2
    g = torch.Generator().manual_seed(23)
    x = (torch.randn(T, H, generator=g) * (1 + 3 * torch.rand(H, generator=g))).to(torch.bfloat16)
    w1 = (torch.randn(E, I, H, generator=g) * 0.05).to(torch.bfloat16)
    w3 = (torch.randn(E, I, H, generator=g) * 0.05).to(torch.bfloat16)
    w2 = (torch.randn(E, H, I, generator=g) * 0.05).to(torch.bfloat16)
    logits = torch.randn(T, E, generator=g)
    logits[:, 4] = -1e9                                           # expert 4 is never routed
    p = torch.softmax(logits, dim=-1)
    topk_w, topk_idx = torch.topk(p, K, dim=-1)
    topk_w = (topk_w / topk_w.sum(-1, keepdim=True))
    geom = O.Geom(O.GROUP, 32)
    weights = [w for e in range(E) for w in (w1[e], w3[e])]
    s_ref, r_ref, l_ref = R.compute_best_scale([x], weights, R.moe_block_parent([w2[e] for e in range(E)], topk_idx, topk_w), geom, O.INT, 4, True)
    for budget in (16 << 30, 3 * 2 * E * I * H * 2):   # all ratios at once / three per chunk
        s, r, l = awq.search_moe_block_mapping(x.cuda(), w1.cuda(), w3.cuda(), w2.cuda(), topk_idx.cuda(), topk_w.cuda(), Args("int4_g32_sym"),
                                               max_variant_bytes=budget)
        assert max(abs(a - b) / b for a, b in zip(l, l_ref)) < 1e-3, [abs(a - b) / b for a, b in zip(l, l_ref)]
        assert r == r_ref
        assert torch.allclose(s, s_ref, rtol=1e-5)


def test_routed_moe_expert_ranges_chain_to_the_single_gpu_output():
    """The expert-parallel layer-wide mapping (awq.search_moe_block_mapping_ep) on one device: ranks own ascending expert ranges, rank r
    continues rank r - 1's running bf16 output with b200q_moe_combine_acc.  Chaining the ranges (2, 3 and 6 "ranks", one range holding a
    never-routed expert, one rank possibly holding no routed pair of some token) must reproduce the single-GPU block output bit for bit;
    summing per-range outputs in fp32 and rounding once -- what a reduce-scatter would do -- must not (that is why the ring exists)."""
    from quantizers_b200 import awq

    E, T, H, I, K = 6, 640, 256, 384, 3
    g = torch.Generator().manual_seed(29)
    x = (torch.randn(T, H, generator=g) * (1 + 3 * torch.rand(H, generator=g))).to(torch.bfloat16).cuda()
    w13 = (torch.randn(E, 2 * I, H, generator=g) * 0.05).to(torch.bfloat16).cuda()
    w2 = (torch.randn(E, H, I, generator=g) * 0.05).to(torch.bfloat16).cuda()
    logits = torch.randn(T, E, generator=g)
    logits[:, 4] = -1e9                                           # expert 4 is never routed
    p = torch.softmax(logits, dim=-1)
    topk_w, topk_idx = torch.topk(p, K, dim=-1)
    topk_w, topk_idx = (topk_w / topk_w.sum(-1, keepdim=True)).cuda(), topk_idx.cuda()
    want = awq.RoutedMoE(x, w2, topk_idx, topk_w)(w13)
    for world in (2, 3, 6):
        per = E // world
        running, partial_sum = None, torch.zeros(T, H, dtype=torch.float32, device="cuda")
        for r in range(world):
            e0 = r * per
            part = awq.RoutedMoE(x, w2[e0:e0 + per], topk_idx, topk_w, expert_offset=e0)
            y = part.project(w13[e0:e0 + per].contiguous())
            partial_sum += part.combine(y.clone()).float()
            running = part.combine(y, out=running if running is not None else None, init=running)
        assert_bits_equal(running, want.cpu(), f"ring of {world}")
        if world == 2:
            assert not torch.equal(partial_sum.to(torch.bfloat16), want)


def test_gptq_hessian_accumulation():
    """§8f rank 4: H = H n/(n+t) + (2/(n+t)) X^T X over three batches (one with a row count that is not a multiple of 8) on the
    tcgen05 fp32-accumulate epilogue against the restated fp32 reference.  Floating point: 2e-5 of max|H| (fp32 summation order)."""
    from quantizers_b200 import gptq

    g = torch.Generator().manual_seed(31)
    K = 384
    H = torch.zeros(K, K, dtype=torch.float32, device="cuda")
    H_ref, n, n_ref = torch.zeros(K, K), 0, 0
    for t in (512, 203, 1024):
        x = (torch.randn(1, t, K, generator=g) * (1 + 3 * torch.rand(K, generator=g))).to(torch.bfloat16)
        H, n = gptq.accumulate_hessian(x.cuda(), H, n)
        H_ref, n_ref = R.accumulate_hessian(x, H_ref, n_ref)
    assert n == n_ref == 1739
    err = (H.cpu() - H_ref).abs().max().item() / H_ref.abs().max().item()
    assert err < 2e-5, err
    assert torch.allclose(H.cpu(), H.cpu().t(), rtol=0, atol=2e-5 * H_ref.abs().max().item())
    with pytest.raises(ValueError):
        gptq.accumulate_hessian(torch.zeros(4, K, device="cuda"), H, n)


@pytest.mark.parametrize("group", [16, 32, 64, 128, 256])
def test_wmean_bf16_fast_path_extremes(group):
    """_compute_layer_means on bf16 (bracketed-reciprocal kernel): exact against the oracle with magnitudes from 1e-30 to > 1
    (denominator outside the bracket's safe range -> IEEE repair), zero groups, ragged row counts and a column count that is
    not a multiple of the 256-column tile."""
    from quantizers_b200 import awq

    g = torch.Generator().manual_seed(77)
    rows, cols = 8 * 13 + 5, 256 * 3 + 256 * (group == 256) + group
    w = torch.randn(rows, cols, generator=g) * 0.02
    w[:, : group] *= 300.0          # |max| > 1
    w[3, group: 2 * group] = 0.0    # zero group: 0 / 1e-6
    w[4, :] = 1e-30
    w[5, :8] = 6e-39                # bf16 subnormals
    w[6, ::3] = -0.0
    w = w.to(torch.bfloat16)
    got = awq.compute_layer_means([w.cuda(), w[:7].cuda()], group).cpu()
    ref = O.w_mean([w, w[:7]], group).float()
    assert torch.equal(got, ref), group


@pytest.mark.parametrize("name,geom,sym", [("int4_g32_sym", O.Geom(O.GROUP, 32), True), ("int4_g128_asym", O.Geom(O.GROUP, 128), False),
                                          ("int4_g32_asym", O.Geom(O.GROUP, 32), False), ("int4_g128_sym", O.Geom(O.GROUP, 128), True)])
def test_scaled_fake_quantize_grid_fast_kernel_bit_exact(name, geom, sym):
    """awq_fq_fast.cu (weight rows kept in registers across the ratio grid, packed bf16 chain, bracketed reciprocal of the fp32
    column scale) against the oracle's scale -> observe -> fake-quantize -> unscale chain, every bit incl. the -0.0 of torch.round:
    adversarial weights (zero groups, -0.0 rows, 1e-30 values, outlier columns, ragged rows / column strips) and scale vectors that
    span 1e-3 .. 1e3, plus ratios whose scales fall outside the reciprocal-safe range (exact path)."""
    from quantizers_b200 import awq
    from tests.util import synth_weight

    g = torch.Generator().manual_seed(17)
    for rows, K in ((77, 1280), (130, 384)):                       # 1280 = 5 full strips of 256; 384 = 1.5 strips
        w = synth_weight(rows, K, torch.bfloat16, 5 + rows)
        R_ = 6
        scales = torch.exp(torch.empty(R_, K).uniform_(-6.9, 6.9, generator=g))
        scales[1] = 1.0
        scales[2, :64] = 3e-26                                      # outside [2^-60, 2^60]: exact chain for those chunks
        scales[3, 100:140] = 7e22
        out = torch.empty(R_, rows, K, dtype=torch.bfloat16, device="cuda")
        awq.scaled_fake_quantize_grid(w.cuda(), scales.cuda(), Args(name), out)
        for r in range(R_):
            want = R.scaled_fake_quantize(w, scales[r], geom, O.INT, 4, sym)
            assert_bits_equal(out[r], want, f"{name} rows={rows} K={K} ratio {r}")
        one = awq.scaled_fake_quantize(w.cuda(), scales[4].cuda(), Args(name))
        assert_bits_equal(one, R.scaled_fake_quantize(w, scales[4], geom, O.INT, 4, sym), f"{name} single")


@pytest.mark.parametrize("H,HKV,D,S,B", [(32, 8, 128, 512, 2), (4, 2, 64, 64, 3), (5, 1, 128, 70, 1), (8, 8, 128, 300, 2), (6, 3, 64, 257, 2),
                                         (32, 8, 128, 512, 12), (16, 4, 64, 384, 24)])   # the last two: several items per persistent stream
def test_attention_core_tcgen05_vs_reference(H, HKV, D, S, B):
    """csrc/awq_attn_core.cu (causal GQA attention of every sample on tcgen05: fp32 scores / statistics, bf16 probabilities, fp32
    accumulation) against an fp32 reference of the same batch-1 causal attention and against torch SDPA: within bf16 output
    resolution; full and ragged sequence lengths, MHA / GQA / MQA, both head sizes."""
    import ctypes

    from quantizers_b200 import _lib as L

    g = torch.Generator().manual_seed(11)
    T = B * S
    qkv = (torch.randn(T, (H + 2 * HKV) * D, generator=g) * 1.5).to(torch.bfloat16).cuda()
    q, k, v = qkv.split([H * D, HKV * D, HKV * D], dim=-1)
    q = q.reshape(B, S, H, D).transpose(1, 2).float()
    k = k.reshape(B, S, HKV, D).transpose(1, 2).float().repeat_interleave(H // HKV, dim=1)
    v = v.reshape(B, S, HKV, D).transpose(1, 2).float().repeat_interleave(H // HKV, dim=1)
    att = (q @ k.transpose(-1, -2)) / D ** 0.5
    att = att.masked_fill(torch.ones(S, S, dtype=torch.bool, device="cuda").triu(1), float("-inf"))
    want = (torch.softmax(att, dim=-1) @ v).transpose(1, 2).reshape(T, H * D)
    lib = L.lib()
    out = torch.full((T, H * D), float("nan"), dtype=torch.bfloat16, device="cuda")
    nws = int(lib.b200q_attention_workspace(T, HKV, D, S))
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    before = qkv.clone()
    L.check(lib.b200q_attention_core(L.ptr(qkv), T, H, HKV, D, S, L.ptr(out), L.ptr(ws), nws, L.stream_ptr(qkv.device)))
    torch.cuda.synchronize()
    assert torch.equal(qkv, before)                      # inputs are read only
    assert bool(torch.isfinite(out.float()).all())
    err = (out.float() - want).abs()
    tol = 2.0 ** -7 * want.abs() + 0.02                  # bf16 output + bf16 probabilities
    assert bool((err <= tol).all()), (float(err.max()), float(want.abs().max()))
    sd = torch.nn.functional.scaled_dot_product_attention(q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16), is_causal=True)
    sd = sd.transpose(1, 2).reshape(T, H * D).float()
    assert float((out.float() - sd).abs().max()) <= 0.03 + 2.0 ** -6 * float(sd.abs().max())
