"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print("%-60s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %8d %12.1f %6.1f%%" % (name[:60], n, t / 1e3, 100 * t / max(total, 1)))
print("%-60s %8d %12.1f" % ("TOTAL", sum(v[0] for v in tot.values()), total / 1e3))
