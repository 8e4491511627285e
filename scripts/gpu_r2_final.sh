#!/bin/bash
# Round-2 closing GPU call: full parity suite, smoke, full bench line + reference arm, kernel timeline, ncu launch
# list and ncu --set full captures of the dominant kernels.  gpurun --timeout 3000 -- bash scripts/gpu_r2_final.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
echo "=== pytest gpu"
timeout -k 10 900 python -m pytest tests/ -q -m gpu --timeout 300 2>&1 | tail -30 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "=== bench (full line)"
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "value_long", "e2e", "value_fast", "value_fp32", "config3_burnin40",
              "config2_cnn_iqn", "cpu_baseline", "cuda_torch_baseline", "acting", "clocks"):
        print(k, json.dumps(d.get(k))[:330])
    r = d["roofline"]; print("roofline", r["shape"], r["achieved"], r["frac"], r["us_per_launch"])
    g = d["roofline_gather"]; print("gather", g["achieved"], g["frac"], g["us_per_launch"])
except Exception as e:
    print("bench parse failed", e)
PY
echo "=== bench --impl reference"
timeout -k 10 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -2 gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
echo "=== kernel timeline"
timeout -k 10 300 python scripts/kernel_trace.py --size 65536 --steps 5 --timeline > gpurun_out/kernel_timeline.txt 2>&1
grep "updates " gpurun_out/kernel_timeline.txt
echo "=== per-GEMM device times"
timeout -k 10 200 python scripts/gemm_profile.py > gpurun_out/gemm_profile.txt 2>&1; tail -1 gpurun_out/gemm_profile.txt | cut -c1-200
echo "=== ncu launch list"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 3000 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --long-steps 2 --size 65536 --no-cpu-baseline --no-side-lines > gpurun_out/ncu_bench.log 2>&1
if [ $(wc -l < gpurun_out/launches.csv) -lt 500 ]; then
  echo "few kernels seen through the graphs: launch list with RT_GRAPHS=0"
  RT_GRAPHS=0 timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 3000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --long-steps 2 --size 65536 --no-cpu-baseline --no-side-lines > gpurun_out/ncu_bench.log 2>&1
fi
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt | head -25
cap() {  # name regex skip count
  timeout -k 10 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$2 -s $3 -c $4 -f -o gpurun_out/prof_$1 \
    python bench.py --steps 1 --warmup 3 --long-steps 1 --size 65536 --no-cpu-baseline --no-side-lines > gpurun_out/ncu_full_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py > gpurun_out/prof_$1_summary.txt
  grep -i "kernel name\|gpu__time_duration\|dram__bytes\|tensor_cycles\|lts__throughput\|dram_throughput" gpurun_out/prof_$1_summary.txt | cut -c1-150
}
echo "=== ncu full: hidden-layer GEMM (persistent tcgen05, BN=256)"
cap gemm "k_gemm_tc_pILi256ELi0ELi0" 3 2
echo "=== ncu full: gather"
cap gather k_gather 3 1
echo "=== ncu full: LSTM recurrence (mma.sync) and one-launch BPTT"
cap lstm k_lstm_seq_mma 1 1
cap bptt k_lstm_bptt_p 1 1
