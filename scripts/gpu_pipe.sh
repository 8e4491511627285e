#!/bin/bash
# Quick GPU check: replay + learner/trainer parity tests, bench with both gather kernels.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_replay_gpu.py tests/test_replay_props_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_pipe.log
for v in 0 1; do
RT_GATHER_BULK=$v timeout -k 10 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b$v.json 2> gpurun_out/bench_b$v.err
tail -3 gpurun_out/bench_b$v.err; cut -c1-200 gpurun_out/bench_b$v.json; echo; grep -o '"e2e": {[^}]*}' gpurun_out/bench_b$v.json; grep -o '"roofline_gather": {[^}]*}' gpurun_out/bench_b$v.json
done
cp gpurun_out/bench_b1.json gpurun_out/bench.json
