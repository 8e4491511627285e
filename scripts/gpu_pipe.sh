#!/bin/bash
# Quick GPU A/B.
mkdir -p gpurun_out
for v in 1 2; do
RT_CONV_SHALLOW=$v timeout -k 10 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s$v.json 2> gpurun_out/bench_s$v.err
tail -3 gpurun_out/bench_s$v.err; echo "shallow=$v"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_s$v.json
done
RT_CONV_SHALLOW=2 timeout -k 10 600 python -m pytest tests/test_trainer_gpu.py tests/test_learner_gpu.py -m gpu -q -x 2>&1 | tail -2
