#!/bin/bash
# Round-2 GPU call 7: BPTT with a 17 KB tile (co-resident with the side GEMMs), early quantile embeddings.
mkdir -p gpurun_out
echo "=== learner / trainer suite"
timeout -k 10 900 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py tests/test_bptt_gpu.py -q -m gpu --timeout 600 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/pytest_default.log
for v in 0 1; do
  echo "=== bench RT_PHI_EARLY=$v"
  RT_PHI_EARLY=$v timeout -k 5 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_phi$v.json 2> gpurun_out/bench_phi$v.err
  tail -2 gpurun_out/bench_phi$v.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_phi$v.json | head -2; grep -o '"e2e": {[^}]*}' gpurun_out/bench_phi$v.json
done
echo "=== timeline"
timeout -k 10 300 python scripts/kernel_trace.py --size 65536 --steps 5 --timeline > gpurun_out/kernel_timeline.txt 2>&1
grep "updates " gpurun_out/kernel_timeline.txt
grep -A100 "timeline of the last update" gpurun_out/kernel_timeline.txt | grep "lstm\|bptt\|k_gemm_tc<128, 1, 1\|k_cos\|gemm_tc_p<128, 0, 0\|quantile_mul \|k_adam\|k_uniform" | cut -c1-100
