"""GPU: registry-named observers against the restated oracle (oracle/llmc_restated.py on the pinned C arithmetic)."""
import pytest
import torch

from oracle import llmc_restated as R
from oracle import oracle as O
from tests.test_gpu_compress import Args
from tests.util import assert_bits_equal, synth_weight

pytestmark = pytest.mark.gpu


def test_registry_names():
    from quantizers_b200.observers import Observer

    assert {"memoryless_minmax", "static_minmax", "minmax"} <= set(Observer.registered_names())
    with pytest.raises(KeyError):
        Observer.load_from_registry("nope", base_name="weight", args=Args("int4_g128_asym"))


@pytest.mark.parametrize("name,geom,qtype,nb,sym", [("int4_g128_asym", O.Geom(O.GROUP, 128), O.INT, 4, False),
                                                    ("int4_g32_sym", O.Geom(O.GROUP, 32), O.INT, 4, True),
                                                    ("fp8_channel", O.Geom(O.CHANNEL, 0), O.FP8, 8, True),
                                                    ("fp8_block", O.Geom(O.BLOCK, 0, 128, 128), O.FP8, 8, True)])
def test_weight_observer_forward_matches_oracle(name, geom, qtype, nb, sym):
    from quantizers_b200.observers import Observer

    w = synth_weight(256 + 64, 512, torch.bfloat16, 3)
    obs = Observer.load_from_registry("memoryless_minmax", base_name="weight", args=Args(name))
    scale, zp = obs(w.cuda())
    mn, mx = O.minmax(w, geom)
    s_ref, z_ref = O.calculate_qparams(mn, mx, qtype, nb, sym)
    assert_bits_equal(scale.reshape(s_ref.shape), s_ref, name)
    if qtype == O.INT:
        assert torch.equal(zp.cpu().reshape(z_ref.shape).to(torch.int8), z_ref.to(torch.int8))


@pytest.mark.parametrize("kind", ["memoryless_minmax", "static_minmax", "minmax"])
def test_observer_state_across_batches(kind):
    """Three weight observations in a row: the state (running / EMA / none) follows the restated MinMaxObserver bit for bit."""
    from quantizers_b200.observers import Observer

    geom = O.Geom(O.GROUP, 128)
    obs = Observer.load_from_registry(kind, base_name="weight", args=Args("int4_g128_asym"))
    ref = R.MinMaxObserver(kind)
    for seed in (1, 2, 3):
        w = synth_weight(64, 256, torch.bfloat16, seed, edge=False) * (1.0 + 0.5 * seed)
        w = w.to(torch.bfloat16)
        mn, mx = obs.get_min_max(w.cuda())
        rmn, rmx = ref.update(*O.minmax(w, geom))
        assert_bits_equal(mn.reshape(rmn.shape), rmn, kind)
        assert_bits_equal(mx.reshape(rmx.shape), rmx, kind)


def test_activation_static_global_scale():
    from quantizers_b200.observers import Observer

    g = torch.Generator().manual_seed(8)
    xs = [(torch.randn(1, 64, 256, generator=g) * (i + 1)).to(torch.bfloat16) for i in range(3)]
    obs = Observer.load_from_registry("static_minmax", base_name="input", args=Args("nvfp4"))
    for x in xs:
        gs = obs.get_global_scale(x.cuda())
    want = R.activation_global_scale(xs)
    assert gs.item() == want.item()
    with pytest.raises(ValueError):
        obs(torch.zeros(0, 256, dtype=torch.bfloat16).cuda())
