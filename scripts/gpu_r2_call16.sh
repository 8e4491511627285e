#!/bin/bash
mkdir -p gpurun_out
echo "=== replay + trainer suites (train quota now kept by the library)"
timeout -k 10 500 python -m pytest tests/test_trainer_gpu.py tests/test_replay_gpu.py tests/test_replay_props_gpu.py tests/test_dropin_reference_gpu.py -q -m gpu --timeout 200 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/pytest_default.log
for v in 0 2; do
  echo "=== bench RT_CONV_SHALLOW=$v"
  RT_CONV_SHALLOW=$v timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_cs$v.json 2> gpurun_out/bench_cs$v.err
  tail -2 gpurun_out/bench_cs$v.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_cs$v.json | head -2
done
