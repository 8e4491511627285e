"""Prints the roofline-relevant columns of an `ncu --page raw --csv` dump, one block per kernel."""
import csv
import sys

KEYS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    print("no ncu rows")
    sys.exit(0)
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("-" * 60)
    for k in KEYS:
        for kk in hdr:
            if kk == k or kk.startswith(k):
                print("%-66s %s %s" % (kk, d.get(kk), u.get(kk, "")))
                break
