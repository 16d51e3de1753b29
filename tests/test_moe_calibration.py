"""MoE calibration plumbing (quantizers_b200/moe_calibration.py): per-expert Linears + calibrate-all-experts forward, checked
against transformers' own sparse block (CPU, fp32)."""
import pytest
import torch

from quantizers_b200 import moe_calibration as MC


def _hf_block(E=8, H=64, I=32, k=2, norm=True, seed=0):
    qm = pytest.importorskip("transformers.models.qwen3_moe.modeling_qwen3_moe")
    from transformers import Qwen3MoeConfig

    cfg = Qwen3MoeConfig(hidden_size=H, moe_intermediate_size=I, num_experts=E, num_experts_per_tok=k, norm_topk_prob=norm,
                         num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=2, vocab_size=32)
    torch.manual_seed(seed)
    blk = qm.Qwen3MoeSparseMoeBlock(cfg)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn_like(p) * 0.2)
    return blk


class _Holder(torch.nn.Module):
    def __init__(self, blk):
        super().__init__()
        self.mlp = blk

    def forward(self, x):
        y = self.mlp(x)
        return y[0] if isinstance(y, tuple) else y


@pytest.mark.parametrize("norm", [True, False])
def test_block_output_unchanged(norm):
    blk = _hf_block(norm=norm)
    model = _Holder(blk)
    x = torch.randn(3, 17, 64)
    with torch.no_grad():
        ref = model(x)
        with MC.moe_calibrate_all_experts(model) as names:
            assert names == ["mlp"]
            # the BLOCK is kept (its own forward / router / buffers); only `experts` is linearized
            assert type(model.mlp) is type(blk) and isinstance(model.mlp.experts, MC.LinearizedExperts)
            assert model.mlp.experts.calibrate_all_experts
            got_all = model(x)
        assert not model.mlp.experts.calibrate_all_experts  # linearized for good, sparse again
        got_sparse = model(x)
    torch.testing.assert_close(got_all, ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got_sparse, ref, rtol=1e-5, atol=1e-6)
    lin = [n for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)]
    assert "mlp.experts.0.gate_proj" in lin and "mlp.experts.7.down_proj" in lin and len(lin) == 24


def test_every_expert_sees_every_token():
    model = _Holder(_hf_block(E=8, k=2))
    x = torch.randn(2, 16, 64)
    seen = {}

    def hook(name):
        def f(mod, args):
            seen[name] = seen.get(name, 0) + args[0].shape[0]
        return f

    with torch.no_grad(), MC.moe_calibrate_all_experts(model):
        hs = [m.register_forward_pre_hook(hook(n)) for n, m in model.named_modules() if n.endswith("gate_proj")]
        model(x)
        assert len(seen) == 8 and set(seen.values()) == {32}
        for h in hs:
            h.remove()
    seen.clear()
    with torch.no_grad():
        hs = [m.register_forward_pre_hook(hook(n)) for n, m in model.named_modules() if n.endswith("gate_proj")]
        model(x)
        for h in hs:
            h.remove()
    assert sum(seen.values()) == 32 * 2                      # sparse: each token reaches exactly k experts


def test_linear_router_and_module_list_experts():
    """transformers-4 style block: Linear router, ModuleList of MLP experts, top_k / norm_topk_prob on the block."""
    torch.manual_seed(1)
    E, H, I, k = 4, 32, 16, 2

    class MLP(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w1, self.w3, self.w2 = torch.nn.Linear(H, I, bias=False), torch.nn.Linear(H, I, bias=False), torch.nn.Linear(I, H, bias=False)

        def forward(self, x):
            return self.w2(torch.nn.functional.silu(self.w1(x)) * self.w3(x))

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate = torch.nn.Linear(H, E, bias=False)
            self.experts = torch.nn.ModuleList([MLP() for _ in range(E)])
            self.top_k, self.norm_topk_prob = k, True

        def forward(self, x):
            shp = x.shape
            x = x.reshape(-1, H)
            p = torch.softmax(self.gate(x), dim=-1, dtype=torch.float)
            w, idx = torch.topk(p, k, dim=-1)
            w = (w / w.sum(-1, keepdim=True)).to(x.dtype)
            out = torch.zeros_like(x)
            for t in range(x.shape[0]):
                for j in range(k):
                    out[t] += w[t, j] * self.experts[int(idx[t, j])](x[t:t + 1])[0]
            return out.reshape(shp)

    model = _Holder(Block())
    x = torch.randn(1, 9, H)
    with torch.no_grad():
        ref = model(x)
        with MC.moe_calibrate_all_experts(model):
            got = model(x)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    assert model.mlp.experts[0].w1.weight.shape == (I, H)    # expert modules kept as they were


def test_block_with_buffers_and_shared_experts_is_kept():
    """MiniMax-M2 / GLM style blocks -- a block-level correction-bias buffer used by the routing, a router called with
    extra arguments, shared experts, w1/w3/w2 expert names.  The block must survive: same output, buffer still in the state dict,
    expert Linears visible under the checkpoint's names, every expert sees every token inside the context."""
    torch.manual_seed(3)
    E, H, I, k = 4, 32, 16, 2

    class Expert(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w1, self.w3, self.w2 = torch.nn.Linear(H, I, bias=False), torch.nn.Linear(H, I, bias=False), torch.nn.Linear(I, H, bias=False)

        def forward(self, x):
            return self.w2(torch.nn.functional.silu(self.w1(x)) * self.w3(x))

    class Gate(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.randn(E, H) * 0.3)

        def forward(self, x, bias):                     # router called with the block's buffer
            return torch.sigmoid(torch.nn.functional.linear(x, self.weight)) + bias

    class MiniMaxM2SparseMoeBlock(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate = Gate()
            self.experts = torch.nn.ModuleList([Expert() for _ in range(E)])
            self.shared_experts = Expert()
            self.register_buffer("e_score_correction_bias", torch.randn(E) * 0.1)

        def forward(self, x):
            shp = x.shape
            x = x.reshape(-1, H)
            scores = self.gate(x, self.e_score_correction_bias)
            w, idx = torch.topk(scores, k, dim=-1)
            w = w / w.sum(-1, keepdim=True)
            out = torch.zeros_like(x)
            for e in range(E):
                pos, tok = torch.where((idx == e).t())
                if tok.numel():
                    out.index_add_(0, tok, self.experts[e](x[tok]) * w[tok, pos, None])
            return (out + self.shared_experts(x)).reshape(shp)

    model = _Holder(MiniMaxM2SparseMoeBlock())
    x = torch.randn(2, 11, H)
    seen = {}
    with torch.no_grad():
        ref = model(x)
        with MC.moe_calibrate_all_experts(model) as names:
            assert names == ["mlp"] and type(model.mlp).__name__ == "MiniMaxM2SparseMoeBlock"
            hs = [m.register_forward_pre_hook(lambda mod, a, n=n: seen.__setitem__(n, seen.get(n, 0) + a[0].shape[0]))
                  for n, m in model.named_modules() if n.endswith(".w1") and ".experts." in n]
            got = model(x)
            for h in hs:
                h.remove()
        sparse = model(x)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sparse, ref, rtol=1e-5, atol=1e-6)
    assert set(seen.values()) == {22} and len(seen) == E            # every expert saw all 22 tokens exactly once
    sd = model.state_dict()
    assert "mlp.e_score_correction_bias" in sd and "mlp.shared_experts.w1.weight" in sd
    assert "mlp.experts.3.w2.weight" in sd and not any(".inner." in k2 for k2 in sd)


def test_fused_experts_take_the_architecture_leaf_names():
    blk = _hf_block()
    type(blk).__name__  # Qwen3MoeSparseMoeBlock -> gate_proj / up_proj / down_proj
    assert MC.expert_leaf_names(blk) == ("gate_proj", "up_proj", "down_proj")

    class MixtralSparseMoeBlock(torch.nn.Module):
        pass

    assert MC.expert_leaf_names(MixtralSparseMoeBlock()) == ("w1", "w3", "w2")
    ex = MC.linearize_experts(blk.experts, ("w1", "w3", "w2"))
    assert sorted(n for n, _ in ex[0].named_children() if n.startswith("w")) == ["w1", "w2", "w3"]


def test_unhandled_experts_container_raises():
    class Odd(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate = torch.nn.Linear(4, 2)
            self.experts = torch.nn.Linear(4, 4)           # neither fused 3-D parameters nor a ModuleList

    with pytest.raises(NotImplementedError):
        MC.replace_moe_blocks(_Holder(Odd()))


def test_linearize_rejects_inconsistent_shapes():
    class Bad(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_up_proj = torch.nn.Parameter(torch.zeros(2, 8, 4))
            self.down_proj = torch.nn.Parameter(torch.zeros(2, 4, 5))

    with pytest.raises(ValueError):
        MC.linearize_experts(Bad())
