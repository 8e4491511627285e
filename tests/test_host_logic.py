"""CPU-only tests of the host-side mirror (no CUDA calls)."""
import numpy as np
import pytest

from rltime_b200.history import device_history as dh


def test_flatten_skeleton_rebuild_roundtrip():
    sample = {"x": (np.zeros((4, 3, 3), np.uint8), np.ones(5, np.float32)), "layer0_state": {},
              "layer1_state": {"hx": np.arange(4, dtype=np.float32), "cx": np.zeros(4, np.float32),
                               "initials": np.float32(1.0)}, "layer2_state": {}}
    leaves = dh._flatten(sample)
    assert [p for p, _ in leaves] == [("x", 0), ("x", 1), ("layer1_state", "hx"), ("layer1_state", "cx"),
                                      ("layer1_state", "initials")]
    skel = dh._skeleton(sample)
    rebuilt = dh._rebuild(skel, [v for _, v in leaves])
    assert set(rebuilt) == set(sample) and rebuilt["layer0_state"] == {} and isinstance(rebuilt["x"], tuple)
    np.testing.assert_array_equal(rebuilt["layer1_state"]["hx"], sample["layer1_state"]["hx"])
    info = [dh._Leaf(v) for _, v in leaves]
    assert [l.nbytes for l in info] == [36, 20, 16, 16, 4]
    assert info[-1].shape == ()


def test_extract_gamma_and_rejects_other_discounts():
    g = 0.997
    assert dh.extract_gamma(lambda n, r, po: (g ** n) * r) == g
    with pytest.raises(NotImplementedError):
        dh.extract_gamma(lambda n, r, po: (g ** n) * r + 1e-3)


def test_anneal_value_matches_reference_formula():
    assert dh._anneal_value(0.4, 0.0, True, 1.0) == 0.4
    assert dh._anneal_value(0.4, 0.5, True, 1.0) == 0.4 + (1.0 - 0.4) * 0.5
    assert dh._anneal_value(0.4, 2.0, 0.8) == 0.4 + (0.8 - 0.4) * 1.0
    assert dh._anneal_value(0.4, 0.7, False) == 0.4


def test_init_params_follow_reference_scheme():
    from rltime_b200.init import init_params
    info = [("model.layers.0.layers.0.weight", (8, 4, 3, 3)), ("model.layers.0.layers.0.bias", (8,)),
            ("model.layers.1.lstm_cell.weight_hh", (16, 4)), ("model.layers.1.lstm_cell.bias_ih", (16,))]
    p = init_params(info, 4, seed=0)
    assert float(p["model.layers.0.layers.0.bias"].abs().max()) == 0.0
    assert float(p["model.layers.0.layers.0.weight"].abs().max()) <= (1.0 / 36) ** 0.5
    assert float(p["model.layers.1.lstm_cell.weight_hh"].abs().max()) <= 0.5
    assert 0 < float(p["model.layers.1.lstm_cell.bias_ih"].abs().max()) <= 0.5


@pytest.mark.parametrize("name", ["online_small", "online_async"])
def test_online_buffer_matches_reference_golden(name):
    """The host-side OnlineHistoryBuffer (row a11) against traces of the reference's class."""
    from oracle import scenario as sc
    from rltime_b200.history import OnlineHistoryBuffer
    from tests.util import assert_trace_equal, load_golden
    p = sc.SCENARIOS[name]
    h = OnlineHistoryBuffer(**sc.history_kwargs(p), discount_function=sc.discount_function)
    trace = sc.run_scenario(name, h, lambda hh: None)
    assert_trace_equal(trace, load_golden("replay_%s.npz" % name))
