#!/bin/bash
# Scaling spot check: gpurun --gpus N -- bash scripts/gpu_scale.sh N
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
tail -3 gpurun_out/bench_$N.err | cut -c1-300; cut -c1-260 gpurun_out/bench_$N.json; echo; grep -o '"e2e": {[^}]*}' gpurun_out/bench_$N.json
