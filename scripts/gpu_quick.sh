#!/bin/bash
# Quick GPU check: learner / trainer parity tests, warm kernel timeline, bench.
# Usage: gpurun -- bash scripts/gpu_quick.sh            (add RT_* switches in front to A/B a change)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_quick.log
timeout -k 10 300 python scripts/kernel_trace.py --size 65536 --steps 5 > gpurun_out/kernel_trace.txt 2>&1
head -24 gpurun_out/kernel_trace.txt | cut -c1-120 | tail -21; grep "updates \|disabled" gpurun_out/kernel_trace.txt | cut -c1-230
timeout -k 10 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json; echo; grep -o '"e2e": {[^}]*}' gpurun_out/bench.json
