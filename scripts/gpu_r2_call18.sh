#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do RT_ACT_GRAPH=$v timeout -k 5 120 python scripts/act_trace.py 2>&1 | tail -2; done
ACT_CALLS=2 RT_ACT_GRAPH=0 timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/act_launches.csv python scripts/act_trace.py 2>&1 | tail -2
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/act_launches.csv')) if len(r)>10 and r[0].isdigit()]
# last acting step: take the final 40 launches
for r in rows[-45:]:
    print(r[4][:60], r[-1])
PY
