#!/bin/bash
# A/B of the overlapped gradient all-reduce at N GPUs: gpurun --gpus N -- bash scripts/gpu_dp_ab.sh N
N=${1:-2}
mkdir -p gpurun_out
for v in 0 1; do
RT_DP_OVERLAP=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$v \
  bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dp$v.json 2> gpurun_out/bench_dp$v.err
echo "overlap=$v"; tail -2 gpurun_out/bench_dp$v.err | cut -c1-200; grep -o '"value": [0-9.]*, "unit": "updates/s", "n_gpus": [0-9]*' gpurun_out/bench_dp$v.json | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_dp$v.json
done
NCCL_MAX_CTAS=4 RT_DP_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err
echo "overlap=1 NCCL_MAX_CTAS=4"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_dp2.json
