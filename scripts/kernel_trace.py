"""Warm kernel timeline of the bench update (CUPTI through torch.profiler; no ncu serialisation).

Prints, per kernel name: launches, total / mean device time, share of the update's GPU-busy time;
then the update span, the busy time inside it and the largest idle gaps (with the kernel that
follows each gap).  Usage: python scripts/kernel_trace.py [--size 65536] [--steps 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--timeline", action="store_true",
                    help="also print every kernel of the last profiled update in start order: offset, duration, stream")
    args = ap.parse_args()
    cfg = dict(bench.CFG)
    cfg["size"] = args.size
    cfg["gemm"] = "tf32"
    import random
    random.seed(0)
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    hist, learner, _ = bench.build_device_workload(cfg, device, seed=0, rank=0)
    for _ in range(5):
        bench.one_update(hist, learner, cfg["B"])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            bench.one_update(hist, learner, cfg["B"])
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    if args.timeline:
        def sn(n):
            n = n.replace("(anonymous namespace)::", "").replace("void ", "")
            return n.split("(")[0][-60:]
        tl = sorted(((e.time_range.start, e.time_range.end, sn(e.name), getattr(e, "device_resource_id", -1))
                     for e in evs), key=lambda r: r[0])
        # the last update starts at the last k_uniform (quantile fractions: first kernel of a learner step)
        starts = [i for i, r in enumerate(tl) if r[2].endswith("k_uniform")]
        if starts:
            i0 = starts[-1]
            # include the replay kernels of this update's draw that ran just before
            t0 = tl[i0][0]
            print("timeline of the last update (us from its first learner kernel; stream = CUPTI stream id)")
            for s_, e_, n_, st_ in tl[max(i0 - 12, 0):]:
                print("  %9.1f %8.1f  s%-4s %s" % (s_ - t0, e_ - s_, st_, n_))
    rows = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda r: r[0])
    if not rows:
        print("no device events captured")
        return
    def short_name(n):
        n = n.replace("(anonymous namespace)::", "").replace("void ", "")
        return n.split("(")[0][-70:]
    rows = [(s, e, short_name(n)) for s, e, n in rows]
    per = {}
    for s, e, n in rows:
        short = n
        a = per.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += e - s
    span = rows[-1][1] - rows[0][0]
    busy = 0.0
    cur_s, cur_e = rows[0][0], rows[0][1]
    gaps = []
    for s, e, n in rows[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, n[-60:]))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print("%-72s %8s %10s %8s %6s" % ("kernel", "launches", "total_us", "mean_us", "share"))
    tot = sum(v[1] for v in per.values())
    for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %8d %10.1f %8.2f %5.1f%%" % (k, c, t / args.steps, t / c, 100 * t / tot))
    print("updates %d  span/update %.1f us  busy/update %.1f us  idle/update %.1f us  kernel-sum/update %.1f us  launches/update %.1f" % (
        args.steps, span / args.steps, busy / args.steps, (span - busy) / args.steps, tot / args.steps,
        len(rows) / args.steps))
    bygap = {}
    for g, n in gaps:
        a = bygap.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += g
    print("idle time by the kernel that follows the gap (per update):")
    for k, (c, t) in sorted(bygap.items(), key=lambda kv: -kv[1][1])[:25]:
        print("  %-62s gaps %5d  idle_us %9.1f  mean %6.2f" % (k, c, t / args.steps, t / c))


if __name__ == "__main__":
    main()
