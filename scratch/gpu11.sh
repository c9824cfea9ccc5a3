#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "tiled or resident" > gpurun_out/k_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/k_pytest.log
tail -12 gpurun_out/k_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
PM_GATHER_TILED=1 timeout 300 $B > gpurun_out/k_bench_t384.json 2> gpurun_out/k_bench_t384.err
for v in p q; do
PM_GATHER_TILED=1 PM_LIB=$PWD/scratch/variants/libpmstep_$v.so timeout 300 $B > gpurun_out/k_bench_$v.json 2> gpurun_out/k_bench_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
