#!/bin/bash
# gpurun --gpus 8 -- bash scripts/gpu_r2_scale8.sh : weak scaling at 8 / 4 GPUs (1M replay sharded by env)
mkdir -p gpurun_out
run() {  # N label env...
  N=$1; label=$2; shift; shift
  env "$@" timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_${N}gpu_$label.json 2> gpurun_out/bench_${N}gpu_$label.err
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_${N}gpu_$label.err | tail -3 | cut -c1-300
  echo "N=$N $label: $(grep -o '"value": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -2 | tr '\n' ' ') $(grep -o '"e2e": {[^}]*}' gpurun_out/bench_${N}gpu_$label.json | cut -c1-60)"
}
run 8 lib16 RT_DP_LIB=1
run 8 lib32 RT_DP_LIB=1 RT_NCCL_MAX_CTAS=32
run 4 lib16 RT_DP_LIB=1
timeout -k 10 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_1gpu_samebox8.json 2> /dev/null
echo "N=1: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_1gpu_samebox8.json | head -2 | tr '\n' ' ')"
