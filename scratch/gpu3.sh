#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "fft or fused or potential or incremental or resident" > gpurun_out/c_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c_pytest.log
tail -25 gpurun_out/c_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/c_bench_default.json 2> gpurun_out/c_bench_default.err
PM_FFT_FUSE=0 timeout 300 $B > gpurun_out/c_bench_v2_nofuse.json 2> gpurun_out/c_bench_v2_nofuse.err
PM_FFT_LAG=24 timeout 300 $B > gpurun_out/c_bench_v2_lag24.json 2> gpurun_out/c_bench_v2_lag24.err
PM_FFT_V2=0 PM_FFT_FUSE=0 timeout 300 $B > gpurun_out/c_bench_v1.json 2> gpurun_out/c_bench_v1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['config'].get('sort'), d['config'].get('fft'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
