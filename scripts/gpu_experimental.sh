#!/bin/bash
# Validation of the compiled-but-off experimental paths (rt_bptt.cuh).  Usage:
#   gpurun --timeout 600 -- bash scripts/gpu_experimental.sh
# Every step runs under its own timeout: the one-launch BPTT kernel is a cooperative kernel with grid
# barriers, so a bug shows up as a hang, not as a wrong number.
mkdir -p gpurun_out
echo "=== persistent BPTT vs stepwise (gradients)"
RT_TEST_EXPERIMENTAL=1 timeout -k 5 120 python -m pytest tests/test_experimental_gpu.py -m gpu -q -x --timeout 100 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_experimental.log
if grep -q "passed" gpurun_out/pytest_experimental.log && ! grep -q "failed\|error\|Timeout" gpurun_out/pytest_experimental.log; then
  echo "=== learner / trainer parity suite with RT_BPTT_PERSISTENT=1"
  RT_BPTT_PERSISTENT=1 timeout -k 5 300 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -5 | cut -c1-300
  for v in 0 1; do
    echo "=== bench RT_BPTT_PERSISTENT=$v"
    RT_BPTT_PERSISTENT=$v timeout -k 5 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bptt$v.json 2> gpurun_out/bench_bptt$v.err
    tail -2 gpurun_out/bench_bptt$v.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_bptt$v.json
  done
  RT_BPTT_PERSISTENT=1 timeout -k 5 200 python scripts/kernel_trace.py --size 65536 --steps 5 2>&1 | grep "bptt\|updates "
fi
