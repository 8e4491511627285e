"""Per-shape device time of every GEMM-shaped launch of the bench update (CUDA events, eager)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

cfg = dict(bench.CFG)
cfg["size"] = 65536
cfg["gemm"] = "tf32"
import random  # noqa: E402
random.seed(0)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
hist, learner, _ = bench.build_device_workload(cfg, dev, seed=0, rank=0)
for _ in range(5):
    bench.one_update(hist, learner, cfg["B"])
learner.profile_gemms(True)
steps = 5
for _ in range(steps):
    bench.one_update(hist, learner, cfg["B"])
per = learner.gemm_launches()
shapes = learner.gemm_shapes()
learner.profile_gemms(False)
agg = {}
for (fl, ms), sh in zip(per, shapes):
    a = agg.setdefault(sh, [0, 0.0, fl])
    a[0] += 1
    a[1] += ms
kinds = {0: "gemm", 1: "conv", 2: "convdW", 3: "convdX"}
print("%-8s %8s %6s %7s %3s %3s %6s %9s %9s %7s" % ("kind", "M", "N", "K", "tA", "tB", "n/upd", "us/launch", "us/update", "TF/s"))
tot = 0.0
for sh, (c, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    us = 1e3 * ms / c
    tot += 1e3 * ms / steps
    print("%-8s %8d %6d %7d %3d %3d %6.1f %9.1f %9.1f %7.0f" % (kinds[sh[0]], sh[1], sh[2], sh[3], sh[4], sh[5],
                                                             c / steps, us, 1e3 * ms / steps, fl / us / 1e6))
print("total GEMM-shaped time per update: %.1f us" % tot)
