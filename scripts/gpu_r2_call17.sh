#!/bin/bash
# acting step from a CUDA graph + persistent DevicePolicy buffers
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_trainer_gpu.py -q -m gpu --timeout 200 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_act.log
for v in 1 0; do
  echo "=== bench RT_ACT_GRAPH=$v"
  RT_ACT_GRAPH=$v timeout -k 5 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_act$v.json 2> gpurun_out/bench_act$v.err
  tail -2 gpurun_out/bench_act$v.err; grep -o '"us_per_vector_step": [0-9.]*' gpurun_out/bench_act$v.json; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_act$v.json | head -1
done
