#!/bin/bash
# Round-2 GPU call 1: parity suite with the round-to-nearest TF32 mode as default, precision sweep of every
# GEMM mode (goldens + 50 full-size updates vs the oracle), validation of the one-launch BPTT kernel,
# bench with the unmodified-reference CPU / torch-CUDA arms.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt; free -g >> gpurun_out/smi.txt
echo "=== pytest gpu (continue past failures)"
timeout -k 10 1500 python -m pytest tests/ -q -m gpu --timeout 600 2>&1 | tail -40 | cut -c1-400 | tee gpurun_out/pytest_gpu.log
echo "=== precision sweep"
timeout -k 10 900 python scripts/precision_sweep.py 50 > gpurun_out/precision_sweep.txt 2>&1; grep "SUMMARY\|FAILED\|Error" gpurun_out/precision_sweep.txt | cut -c1-300
echo "=== persistent BPTT vs stepwise (gradients)"
RT_TEST_EXPERIMENTAL=1 timeout -k 5 180 python -m pytest tests/test_experimental_gpu.py -m gpu -q -x --timeout 150 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_experimental.log
if grep -q "passed" gpurun_out/pytest_experimental.log && ! grep -q "failed\|error\|Timeout" gpurun_out/pytest_experimental.log; then
  for v in 0 1; do
    echo "=== bench RT_BPTT_PERSISTENT=$v"
    RT_BPTT_PERSISTENT=$v timeout -k 5 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_bptt$v.json 2> gpurun_out/bench_bptt$v.err
    tail -2 gpurun_out/bench_bptt$v.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_bptt$v.json | head -2
  done
  RT_BPTT_PERSISTENT=1 timeout -k 5 200 python scripts/kernel_trace.py --size 65536 --steps 5 > gpurun_out/kernel_trace_bptt1.txt 2>&1
  grep "bptt\|updates " gpurun_out/kernel_trace_bptt1.txt | cut -c1-200
fi
echo "=== kernel trace (default)"
timeout -k 10 300 python scripts/kernel_trace.py --size 65536 --steps 5 > gpurun_out/kernel_trace.txt 2>&1
head -40 gpurun_out/kernel_trace.txt | cut -c1-140
echo "=== bench (full line)"
timeout -k 10 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "value_long", "e2e", "value_fast", "value_fp32", "config3_burnin40",
              "config2_cnn_iqn", "cpu_baseline", "cuda_torch_baseline", "acting", "clocks"):
        print(k, json.dumps(d.get(k))[:600])
    r = d["roofline"]; print("roofline", r["shape"], r["achieved"], r["frac"], r["us_per_launch"])
except Exception as e:
    print("bench parse failed", e)
PY
echo "=== bench --impl reference"
timeout -k 10 1200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -2 gpurun_out/bench_reference.err; cut -c1-1800 gpurun_out/bench_reference.json
