#!/bin/bash
mkdir -p gpurun_out
echo "=== learner / trainer suite"
timeout -k 10 400 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py tests/test_bptt_gpu.py tests/test_gemm_gpu.py -q -m gpu --timeout 200 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/pytest_default.log
echo "=== bench"
timeout -k 5 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
tail -2 gpurun_out/bench_q.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_q.json | head -2; grep -o '"e2e": {[^}]*}' gpurun_out/bench_q.json
echo "=== timeline"
timeout -k 10 300 python scripts/kernel_trace.py --size 65536 --steps 5 --timeline > gpurun_out/kernel_timeline.txt 2>&1
grep "updates " gpurun_out/kernel_timeline.txt
grep "splitk\|colsum" gpurun_out/kernel_timeline.txt | tail -4 | cut -c1-110
