#!/bin/bash
# ncu launch list of two bench updates (cold-cache, serialised: compare shares, not absolutes)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 3000 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --size 65536 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt | head -50
