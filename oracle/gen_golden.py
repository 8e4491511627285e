"""Generates tests/golden/replay_<scenario>.npz by executing the UNMODIFIED reference
(opherlieber/rltime at /root/reference) in the build container.

TEST INFRASTRUCTURE.  Run:  python oracle/gen_golden.py
Needs /root/reference (read-only) and the gym stub in oracle/stubs; neither exists on
the GPU box, which is why the outputs are committed as fixtures.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("RLTIME_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "stubs"), REF, ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import scenario as sc  # noqa: E402


def make_reference_history(p):
    from rltime.history.prioritized_replay_history import PrioritizedReplayHistoryBuffer
    from rltime.history.replay_history import ReplayHistoryBuffer
    from rltime.general.backend import StateStore
    from rltime.history.online_history import OnlineHistoryBuffer
    cls = {"per": PrioritizedReplayHistoryBuffer, "uniform": ReplayHistoryBuffer,
           "online": OnlineHistoryBuffer}[p["kind"]]
    h = cls(**sc.history_kwargs(p), discount_function=sc.discount_function,
            state_store=StateStore("cpu"))
    if p["kind"] == "per":
        rec = {}
        orig = h._sample_proportional

        def wrapped(batch_size):
            res = orig(batch_size)
            rec["idx"] = list(res)
            return res
        h._sample_proportional = wrapped
        h._rec = rec
    return h


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for name, p in sc.SCENARIOS.items():
        if only and name not in only:
            continue
        h = make_reference_history(p)
        if p["kind"] == "per":
            trace = sc.run_scenario(name, h, lambda hh: hh._rec.get("idx"),
                                    lambda hh: hh._it_sum.sum())
        else:
            trace = sc.run_scenario(name, h, lambda hh: None)
        path = os.path.join(out_dir, "replay_%s.npz" % name)
        np.savez_compressed(path, **trace)
        print("%-18s calls=%d none=%d fields=%d size=%.1f KB" % (
            name, trace["num_calls"], len(trace["none_calls"]), len(trace),
            os.path.getsize(path) / 1024))


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
