"""Pins oracle/replay_oracle.py against golden traces produced by the unmodified
reference (oracle/gen_golden.py).  CPU only."""
import pytest

from oracle import replay_oracle as ro
from oracle import scenario as sc
from tests.util import assert_trace_equal, load_golden


def _make(p):
    cls = {"per": ro.PrioritizedReplayOracle, "uniform": ro.ReplayOracle,
           "online": ro.OnlineOracle}[p["kind"]]
    return cls(**sc.history_kwargs(p), discount_function=sc.discount_function)


@pytest.mark.parametrize("name", sorted(sc.SCENARIOS))
def test_oracle_matches_reference_golden(name):
    p = sc.SCENARIOS[name]
    h = _make(p)
    if p["kind"] == "per":
        trace = sc.run_scenario(name, h, lambda hh: hh.last_sampled_idxes,
                                lambda hh: hh.sum_tree.root())
    else:
        trace = sc.run_scenario(name, h, lambda hh: None)
    want = load_golden("replay_%s.npz" % name)
    # every field bit-exact: indices, frames, fp64 returns / IS weights / tree sums
    assert_trace_equal(trace, want)
