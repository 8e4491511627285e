#!/bin/bash
# Multi-GPU round-2 call: gpurun --gpus N -- bash scripts/gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
if [ "$N" = "2" ]; then
  echo "=== multi-GPU parity tests"
  timeout -k 10 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu --timeout 600 2>&1 | tail -15 | cut -c1-400 | tee gpurun_out/pytest_multigpu.log
fi
run() {  # label, env...
  label=$1; shift
  env "$@" timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_${N}gpu_$label.json 2> gpurun_out/bench_${N}gpu_$label.err
  tail -2 gpurun_out/bench_${N}gpu_$label.err | cut -c1-300
  echo "$label: $(grep -o '"value": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -2 | tr '\n' ' ') $(grep -o '"e2e": {[^}]*}' gpurun_out/bench_${N}gpu_$label.json | cut -c1-60)"
}
run lib16 RT_DP_LIB=1
run lib8 RT_DP_LIB=1 RT_NCCL_MAX_CTAS=8
run lib32 RT_DP_LIB=1 RT_NCCL_MAX_CTAS=32
run torchdist RT_DP_LIB=0
echo "=== single-GPU reference point on the same box"
timeout -k 10 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_1gpu_samebox.json 2> gpurun_out/bench_1gpu_samebox.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_1gpu_samebox.json | head -2
