"""CPU: the reference arm of bench.py really drives the unmodified reference (baseline/_ref) through its own
IQN.train() -- one warm-up + one timed update on a small replay."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_arm  # noqa: E402


@pytest.mark.skipif(not ref_arm.available(), reason="baseline/_ref not installed (baseline/install_ref.sh)")
def test_reference_arm_runs_the_unmodified_reference():
    sys.path.insert(0, ROOT)
    from bench import CFG
    res = ref_arm.run(dict(CFG), device="cpu", threads=4, steps=1, warmup=1, size=4000)
    assert res["updates_timed"] == 1 and res["updates_per_s"] > 0
    assert res["replay_transitions"] >= 4000
    # the reference's own timers (policy_trainer.py:228-244) saw its own phases
    assert {"get_train_data", "calc_target_values", "train"} <= set(res["timings_mean_ms"])
    import rltime
    assert os.path.realpath(rltime.__file__).startswith(os.path.realpath(ref_arm.REF_DIR))
