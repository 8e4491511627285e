"""CUDA history buffers vs the reference's golden traces and vs the oracle (GPU box)."""
import numpy as np
import pytest

from oracle import scenario as sc
from tests.util import assert_trace_equal, load_golden


def make_device_history(p, **kw):
    from rltime_b200.history import (DevicePrioritizedReplayHistoryBuffer,
                                     DeviceReplayHistoryBuffer)
    cls = DevicePrioritizedReplayHistoryBuffer if p["kind"] == "per" else DeviceReplayHistoryBuffer
    return cls(**sc.history_kwargs(p), discount_function=sc.discount_function,
               max_envs=p["envs"], **kw)


def run_device_scenario(name, **kw):
    p = sc.SCENARIOS[name]
    h = make_device_history(p, **kw)
    try:
        if p["kind"] == "per":
            return sc.run_scenario(name, h, lambda hh: hh.last_sampled_idxes,
                                   lambda hh: hh.tree_sum())
        return sc.run_scenario(name, h, lambda hh: None)
    finally:
        h.close()


def check_against(trace, want):
    # bit-exact: sampled indices, loss indices, n-steps, masks, actions, every state byte,
    # fp64 n-step returns and the fp64 tree sum.  The IS weights go through CUDA's pow()
    # (not bit-equal to glibc): fp64 relative tolerance 1e-13, stated here.
    assert_trace_equal(trace, want, skip=("extra_data/importance_weights",))
    if "extra_data/importance_weights" in want:
        np.testing.assert_allclose(trace["extra_data/importance_weights"],
                                   want["extra_data/importance_weights"], rtol=1e-13, atol=0)


DEVICE_SCENARIOS = sorted(k for k, v in sc.SCENARIOS.items() if v["kind"] != "online")


@pytest.mark.gpu
@pytest.mark.parametrize("name", DEVICE_SCENARIOS)
def test_device_matches_reference_golden(name):
    check_against(run_device_scenario(name), load_golden("replay_%s.npz" % name))


@pytest.mark.gpu
def test_numpy_output_mode_matches():
    trace = run_device_scenario("per_async", output="numpy")
    check_against(trace, load_golden("replay_per_async.npz"))
