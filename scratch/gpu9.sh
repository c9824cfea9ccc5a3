#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/i_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/i_pytest.log
tail -12 gpurun_out/i_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/i_bench_tiled.json 2> gpurun_out/i_bench_tiled.err
PM_GATHER_TILED=0 timeout 300 $B > gpurun_out/i_bench_flat.json 2> gpurun_out/i_bench_flat.err
for v in m n; do
PM_LIB=$PWD/scratch/variants/libpmstep_$v.so timeout 300 $B > gpurun_out/i_bench_$v.json 2> gpurun_out/i_bench_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/i_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
