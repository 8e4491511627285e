#!/bin/bash
# Round-2 GPU call 3: full parity suite with the one-launch BPTT and the paired conv1 as defaults, A/B of the
# paired conv1, kernel trace, full bench line.
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout -k 10 1800 python -m pytest tests/ -q -m gpu --timeout 600 2>&1 | tail -40 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
for v in 0 1; do
  echo "=== bench RT_CONV1_PAIR=$v"
  RT_CONV1_PAIR=$v timeout -k 5 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_pair$v.json 2> gpurun_out/bench_pair$v.err
  tail -2 gpurun_out/bench_pair$v.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_pair$v.json | head -2; grep -o '"e2e": {[^}]*}' gpurun_out/bench_pair$v.json
done
echo "=== kernel trace"
timeout -k 10 300 python scripts/kernel_trace.py --size 65536 --steps 5 > gpurun_out/kernel_trace.txt 2>&1
head -30 gpurun_out/kernel_trace.txt | cut -c1-130; grep "updates " gpurun_out/kernel_trace.txt
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "=== bench (full line)"
timeout -k 10 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "value_long", "e2e", "value_fast", "value_fp32", "config3_burnin40",
              "config2_cnn_iqn", "cpu_baseline", "cuda_torch_baseline", "acting", "clocks"):
        print(k, json.dumps(d.get(k))[:400])
    r = d["roofline"]; print("roofline", r["shape"], r["achieved"], r["frac"], r["us_per_launch"])
except Exception as e:
    print("bench parse failed", e)
PY
