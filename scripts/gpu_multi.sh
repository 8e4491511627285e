#!/bin/bash
# Multi-GPU round: gpurun --gpus N -- bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "=== dist_check ($N ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  scripts/dist_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -15 | tee gpurun_out/dist_check_$N.log
echo "=== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
tail -3 gpurun_out/bench_$N.err | cut -c1-300; cut -c1-300 gpurun_out/bench_$N.json; echo; grep -o '"e2e": {[^}]*}' gpurun_out/bench_$N.json
echo "=== bench --gpus 1 (same box)"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
cut -c1-300 gpurun_out/bench_1.json; echo; grep -o '"e2e": {[^}]*}' gpurun_out/bench_1.json
