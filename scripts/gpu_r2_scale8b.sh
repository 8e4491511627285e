#!/bin/bash
mkdir -p gpurun_out
run() {  # N label env...
  N=$1; label=$2; shift; shift
  env "$@" timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-side-lines > gpurun_out/bench_${N}gpu_$label.json 2> gpurun_out/bench_${N}gpu_$label.err
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_${N}gpu_$label.err | tail -2 | cut -c1-300
  echo "N=$N $label: $(grep -o '"value": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${N}gpu_$label.json | head -2 | tr '\n' ' ')"
}
run 8 hi_c16 RT_DP_COMM_PRIO=1
run 8 lo_c16 RT_DP_COMM_PRIO=0
run 8 hi_c32 RT_DP_COMM_PRIO=1 RT_NCCL_MAX_CTAS=32 RT_DP_RESERVE_SMS=32
run 8 hi_c32r0 RT_DP_COMM_PRIO=1 RT_NCCL_MAX_CTAS=32 RT_DP_RESERVE_SMS=0
run 4 hi_c16 RT_DP_COMM_PRIO=1
