"""Times the batch gather alone (CUDA events around the kernel) for the register-staged kernel and
the bulk-copy ring geometries."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import random  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

cfg = dict(bench.CFG)
cfg["size"] = 131072
cfg["gemm"] = "tf32"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
S, n, B = cfg["T"], cfg["n"], cfg["B"]
nbytes = 2 * (S + n) * B * (bench.FRAME_BYTES + 2 * cfg["units"] * 4 + 4)
variants = [("register-staged k_gather", {"RT_GATHER_BULK": "0"})]
for c, name in enumerate(["8K x8 look6", "16K x6 look4", "4K x16 look12", "32K x4 look2"]):
    for cps in (1, 2, 3):
        if (c, cps) in ((1, 3), (3, 2), (3, 3)):
            continue            # shared memory: 96 KB / 128 KB rings
        variants.append(("bulk %s, %d CTA/SM" % (name, cps), {"RT_GATHER_BULK": "1", "RT_GB_CFG": str(c), "RT_GB_CPS": str(cps)}))
from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer  # noqa: E402
for name, env in variants:
    os.environ.update(env)
    random.seed(0)
    hist, learner, _ = bench.build_device_workload(cfg, dev, seed=0, rank=0)
    for _ in range(5):
        hist.get_train_data(B, 0.0)
    torch.cuda.synchronize()
    hist.profile_gather(True)
    for _ in range(100):
        hist.get_train_data(B, 0.0)
        torch.cuda.synchronize()
    ms, k = hist.gather_time()
    us = 1e3 * ms / k
    print("%-34s %6.2f us/launch  %7.0f GB/s" % (name, us, nbytes / us / 1e3))
    hist.close()
    learner.close()
    del hist, learner
    torch.cuda.empty_cache()
