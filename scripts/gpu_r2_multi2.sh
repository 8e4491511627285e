#!/bin/bash
# gpurun --gpus 2 -- bash scripts/gpu_r2_multi2.sh : DP backward as one graph (external event), A/B vs the split schedule
N=2
mkdir -p gpurun_out
echo "=== multi-GPU parity tests"
timeout -k 10 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu --timeout 600 2>&1 | tail -15 | cut -c1-400 | tee gpurun_out/pytest_multigpu.log
run() {  # label, env...
  label=$1; shift
  env "$@" timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_${N}gpu_$label.json 2> gpurun_out/bench_${N}gpu_$label.err
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_${N}gpu_$label.err | tail -2 | cut -c1-300
  echo "$label: $(grep -o '"value": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -2 | tr '\n' ' ') $(grep -o '"e2e": {[^}]*}' gpurun_out/bench_${N}gpu_$label.json | cut -c1-60)"
}
run event16 RT_DP_LIB=1
run event32 RT_DP_LIB=1 RT_NCCL_MAX_CTAS=32
run event8 RT_DP_LIB=1 RT_NCCL_MAX_CTAS=8
run split16 RT_DP_LIB=1 RT_DP_SPLIT=1
run torchdist RT_DP_LIB=0
echo "=== single-GPU reference point + timeline on the same box"
timeout -k 10 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_1gpu_samebox.json 2> gpurun_out/bench_1gpu_samebox.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_1gpu_samebox.json | head -2
timeout -k 10 300 python scripts/kernel_trace.py --size 65536 --steps 5 --timeline > gpurun_out/kernel_timeline.txt 2>&1
grep "updates " gpurun_out/kernel_timeline.txt
