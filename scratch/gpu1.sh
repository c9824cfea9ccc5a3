#!/bin/bash
# session A: correctness of the incremental sort + fused plane FFT, then A/B bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 -x -k "incremental or fused_plane" > gpurun_out/a_pytest_new.log 2>&1
echo "new tests rc=$?" >> gpurun_out/a_pytest_new.log
tail -5 gpurun_out/a_pytest_new.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
PM_SORT=full timeout 300 $B > gpurun_out/a_bench_sortfull.json 2> gpurun_out/a_bench_sortfull.err
PM_FFT_FUSE=0 timeout 300 $B > gpurun_out/a_bench_nofuse.json 2> gpurun_out/a_bench_nofuse.err
PM_FFT_LAG=6 timeout 300 $B > gpurun_out/a_bench_lag6.json 2> gpurun_out/a_bench_lag6.err
PM_FFT_LAG=24 timeout 300 $B > gpurun_out/a_bench_lag24.json 2> gpurun_out/a_bench_lag24.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['config'].get('sort'), d['config'].get('fft'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/a_pytest_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/a_pytest_all.log
tail -8 gpurun_out/a_pytest_all.log
