#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "tiled" > gpurun_out/m_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/m_pytest.log
tail -5 gpurun_out/m_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
PM_GATHER_TILED=1 timeout 300 $B > gpurun_out/m_bench_t.json 2> gpurun_out/m_bench_t.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/m_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
PM_GATHER_TILED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gather_tiled" -s 4 -c 1 -f -o gpurun_out/m_prof_gt python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/m_ncu.log 2>&1
ls -la gpurun_out/m_prof_gt.ncu-rep
