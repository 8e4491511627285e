"""Precision evidence (profiles/r02_precision_*.txt): every GEMM mode of the library against
(a) the reference goldens, every update of every case, and (b) the torch fp32 oracle over N
consecutive full-size updates.  Usage: python scripts/precision_sweep.py [updates]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.learner_cases import CASES  # noqa: E402
from tests.test_learner_gpu import full_size_drift, golden_errors  # noqa: E402

MODES = ["fp32", "tf32_trunc", "tf32"]

if __name__ == "__main__":
    updates = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    print("== goldens: max |d| to the reference golden per update (targets / qloss / td_mean / report)")
    for name in sorted(CASES):
        for mode in MODES:
            try:
                rows = golden_errors(name, mode)
            except Exception as ex:  # noqa: BLE001
                print("%-22s %-10s FAILED %s" % (name, mode, str(ex)[:200]))
                continue
            print("%-22s %-10s " % (name, mode) + " | ".join(
                "u%d t %.1e q %.1e m %.1e r %.1e" % (u, e["targets"], e["qloss"], e["td_mean"], e["report"])
                for u, e in enumerate(rows)))
    for tf in (True, False):
        print("== full size, %d consecutive updates vs the torch fp32 oracle, %s" % (
            updates, "weights reset to the oracle's before every update (teacher-forced)" if tf else "free-running"))
        res = full_size_drift(MODES, updates, log=print, teacher_forced=tf)
        for mode, rows in res.items():
            print("SUMMARY %s %-10s max|d| qloss %.2e td_mean %.2e report %.2e targets %.2e (oracle |targets| max %.3f) "
                  "near-tie rows %d flipped rows %d" % (
                      "teacher-forced" if tf else "free-running", mode, max(e["qloss"] for e in rows),
                      max(e["td_mean"] for e in rows), max(e["report"] for e in rows),
                      max(e["targets"] for e in rows), max(e["ref_targets_absmax"] for e in rows),
                      sum(e["tie_rows"] for e in rows), sum(e["flipped_rows"] for e in rows)))
