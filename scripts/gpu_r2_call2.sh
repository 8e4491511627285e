#!/bin/bash
# Round-2 GPU call 2: full parity suite (new model families, trainer surface, drop-in tests with the unmodified
# reference), precision sweep (teacher-forced + free-running), one-launch BPTT v2 validation and timing.
mkdir -p gpurun_out
echo "=== pytest gpu (continue past failures)"
timeout -k 10 1800 python -m pytest tests/ -q -m gpu --timeout 600 2>&1 | tail -60 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
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
  echo "=== learner suite with RT_BPTT_PERSISTENT=1"
  RT_BPTT_PERSISTENT=1 timeout -k 5 600 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py -m gpu -q --timeout 300 2>&1 | tail -5 | cut -c1-300
fi
echo "=== precision sweep"
timeout -k 10 1200 python scripts/precision_sweep.py 50 > gpurun_out/precision_sweep.txt 2>&1; grep "SUMMARY\|FAILED\|Error" gpurun_out/precision_sweep.txt | cut -c1-300
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
