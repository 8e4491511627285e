#!/bin/bash
# Round-2 GPU call 4: validation + timing of the mma.sync LSTM recurrence (RT_LSTM_MMA=1), the fused trunk ReLU
# derivative and the frame-conversion prefetch; full suite; bench.
mkdir -p gpurun_out
echo "=== learner / trainer suite, default"
timeout -k 10 900 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py tests/test_bptt_gpu.py -q -m gpu --timeout 600 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/pytest_default.log
echo "=== learner / trainer suite, RT_LSTM_MMA=1"
RT_LSTM_MMA=1 timeout -k 10 900 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py tests/test_bptt_gpu.py -q -m gpu --timeout 300 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/pytest_lstm_mma.log
for v in 0 1; do
  echo "=== bench RT_LSTM_MMA=$v"
  RT_LSTM_MMA=$v timeout -k 5 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_mma$v.json 2> gpurun_out/bench_mma$v.err
  tail -2 gpurun_out/bench_mma$v.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_mma$v.json | head -2; grep -o '"e2e": {[^}]*}' gpurun_out/bench_mma$v.json
done
echo "=== kernel trace RT_LSTM_MMA=1"
RT_LSTM_MMA=1 timeout -k 10 300 python scripts/kernel_trace.py --size 65536 --steps 5 > gpurun_out/kernel_trace_mma1.txt 2>&1
head -14 gpurun_out/kernel_trace_mma1.txt | cut -c1-130; grep "lstm\|updates " gpurun_out/kernel_trace_mma1.txt | cut -c1-160
echo "=== full gpu suite"
timeout -k 10 1800 python -m pytest tests/ -q -m gpu --timeout 600 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
