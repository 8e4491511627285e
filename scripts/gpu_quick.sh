#!/bin/bash
# Quick GPU check: learner parity tests, LSTM timeline, bench.  Usage: gpurun -- bash scripts/gpu_quick.sh
mkdir -p gpurun_out
echo "=== learner tests"
timeout -k 10 600 python -m pytest tests/test_learner_gpu.py -m gpu -q -x --timeout 200 -s 2>&1 | tail -40 | cut -c1-300 | tee gpurun_out/pytest_learner.log
echo "=== timeline"
timeout -k 10 120 python scripts/timeline.py 2>&1 | tail -24 | tee gpurun_out/lstm_timeline.txt
echo "=== bench"
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cut -c1-700 gpurun_out/bench.json
if [ -n "$1" ]; then
echo "=== bench with RT_LSTM_UPC=8"
RT_LSTM_UPC=8 timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_upc8.json 2> gpurun_out/bench_upc8.err
tail -3 gpurun_out/bench_upc8.err; cut -c1-400 gpurun_out/bench_upc8.json
RT_LSTM_UPC=8 timeout -k 10 120 python scripts/timeline.py 2>&1 | tail -24 | tee gpurun_out/lstm_timeline_upc8.txt
fi
