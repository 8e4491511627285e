#!/bin/bash
# Closing check after the CTA-pair GEMM: full GPU suite, short bench, ncu --set full of the pair kernel, warm trace.
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout -k 10 330 python -m pytest tests/ -q -m gpu --timeout 200 -x 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "=== bench (no baselines / side lines)"
timeout -k 10 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err
tail -2 gpurun_out/bench_pair.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_pair.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "value_long", "e2e", "acting", "clocks", "gpu_launches"):
        print(k, json.dumps(d.get(k))[:300])
    r = d["roofline"]; print("roofline", r.get("shape"), r["achieved"], r["frac"], r.get("us_per_launch"))
except Exception as e:
    print("bench parse failed", e)
PY
echo "=== ncu full: CTA-pair hidden-layer GEMM"
timeout -k 10 120 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_gemm_tc_pair -s 3 -c 2 -f -o gpurun_out/prof_gemm_pair \
  python bench.py --steps 1 --warmup 3 --long-steps 1 --size 65536 --no-cpu-baseline --no-side-lines > gpurun_out/ncu_full_gemm_pair.log 2>&1
ncu -i gpurun_out/prof_gemm_pair.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py > gpurun_out/prof_gemm_pair_summary.txt
grep -i "kernel name\|gpu__time_duration\|dram__bytes\|tensor_cycles\|lts__throughput" gpurun_out/prof_gemm_pair_summary.txt | cut -c1-150
echo "=== kernel timeline"
timeout -k 10 100 python scripts/kernel_trace.py --size 65536 --steps 5 --timeline > gpurun_out/kernel_timeline_pair.txt 2>&1
grep "updates \|k_gemm_tc_pair" gpurun_out/kernel_timeline_pair.txt | tail -4
