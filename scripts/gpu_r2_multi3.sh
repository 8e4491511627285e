#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_${N}gpu_$label.json 2> gpurun_out/bench_${N}gpu_$label.err
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_${N}gpu_$label.err | tail -2 | cut -c1-300
  echo "N=$N $label: $(grep -o '"value": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -2 | tr '\n' ' ')"
}
run res0 RT_DP_RESERVE_SMS=0
run res16 RT_DP_RESERVE_SMS=16
run res32c32 RT_DP_RESERVE_SMS=32 RT_NCCL_MAX_CTAS=32
run res8c8 RT_DP_RESERVE_SMS=8 RT_NCCL_MAX_CTAS=8
timeout -k 10 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_1gpu_samebox.json 2> /dev/null
echo "N=1: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_1gpu_samebox.json | head -2 | tr '\n' ' ')"
