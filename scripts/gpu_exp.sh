#!/bin/bash
# Experiment runner: LSTM tests + timeline under several env settings.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_learner_gpu.py -m gpu -q -x --timeout 200 2>&1 | tail -5 | cut -c1-300
for cfg in "RT_LSTM_EXP=0" "RT_LSTM_EXP=1" "RT_LSTM_EXP=2"; do
  echo "=== $cfg"
  env $cfg timeout -k 10 120 python scripts/timeline.py 2>&1 | tail -16 | cut -c1-250
done 2>&1 | tee gpurun_out/exp.log
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
