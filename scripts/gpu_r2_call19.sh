#!/bin/bash
# ncu --set full of the out-layer / dueling kernel (3 x 34 us per update)
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_heads_out -s 3 -c 2 -f -o gpurun_out/prof_heads \
  python bench.py --steps 1 --warmup 3 --long-steps 1 --size 65536 --no-cpu-baseline --no-side-lines > gpurun_out/ncu_full_heads.log 2>&1
ncu -i gpurun_out/prof_heads.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/prof_heads_raw.csv
python scripts/ncu_pick.py < gpurun_out/prof_heads_raw.csv > gpurun_out/prof_heads_summary.txt
cat gpurun_out/prof_heads_summary.txt | cut -c1-150
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_heads_raw.csv')))
hdr=rows[0]
for r in rows[2:3]:
    d=dict(zip(hdr,r))
    st=[(k,d[k]) for k in hdr if 'issue_stalled' in k and 'not_issued' not in k and k.endswith('ratio')]
    def f(x):
        try: return float(x.replace(',',''))
        except: return 0
    for k,v in sorted(st,key=lambda kv:-f(kv[1]))[:8]: print(k,v)
    for k in hdr:
        if any(s in k for s in ('l1tex__t_sector_hit_rate','lts__t_sector_hit_rate','achieved_occupancy','sm__warps_active','dram__throughput','l1tex__throughput','lts__t_sectors_srcunit_tex_op_read.sum','smsp__inst_executed.sum','l1tex__data_pipe')):
            print(k,d[k])
PY
