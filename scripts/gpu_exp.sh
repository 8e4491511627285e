#!/bin/bash
mkdir -p gpurun_out
for cfg in "RT_CONV_PERSISTENT=0" "RT_CONV_PERSISTENT=1"; do
  echo "=== $cfg"
  env $cfg timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
done 2>&1 | tee gpurun_out/exp.log
RT_CONV_PERSISTENT=1 timeout -k 10 600 python -m pytest tests/test_learner_gpu.py -m gpu -q -x --timeout 200 2>&1 | tail -3 | cut -c1-300
