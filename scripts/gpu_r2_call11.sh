#!/bin/bash
mkdir -p gpurun_out
echo "=== learner / trainer suite"
timeout -k 10 420 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py tests/test_bptt_gpu.py -q -m gpu --timeout 200 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/pytest_default.log
for v in 0 1; do
  echo "=== bench RT_HEADS_FUSED=$v"
  RT_HEADS_FUSED=$v timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_hf$v.json 2> gpurun_out/bench_hf$v.err
  tail -2 gpurun_out/bench_hf$v.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_hf$v.json | head -2; grep -o '"e2e": {[^}]*}' gpurun_out/bench_hf$v.json
done
echo "=== timeline"
timeout -k 10 200 python scripts/kernel_trace.py --size 65536 --steps 5 --timeline > gpurun_out/kernel_timeline.txt 2>&1
grep "updates " gpurun_out/kernel_timeline.txt
sed -n '/^kernel /,/^updates/p' gpurun_out/kernel_timeline.txt | head -12 | cut -c1-110
