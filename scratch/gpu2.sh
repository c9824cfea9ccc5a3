#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 150 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_bench.log 2>&1
python - <<'PY'
import csv,re
rows=[]
for r in csv.reader(open('gpurun_out/b_launches.csv')):
    if len(r)>10 and r[0].isdigit(): rows.append(r)
hdr=None
for r in csv.reader(open('gpurun_out/b_launches.csv')):
    if 'Kernel Name' in r: hdr=r;break
if hdr is None: raise SystemExit(open('gpurun_out/b_launches.csv').read()[:2000])
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[:70]: print(r[ki][:70], r[vi])
PY
