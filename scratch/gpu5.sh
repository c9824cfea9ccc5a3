#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/e_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/e_pytest.log
tail -15 gpurun_out/e_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/e_bench_default.json 2> gpurun_out/e_bench_default.err
PM_FFT_V3=2 timeout 300 $B > gpurun_out/e_bench_v3_2.json 2> gpurun_out/e_bench_v3_2.err
PM_FFT_V3=3 timeout 300 $B > gpurun_out/e_bench_v3_3.json 2> gpurun_out/e_bench_v3_3.err
PM_FFT_ZMIX=0 timeout 300 $B > gpurun_out/e_bench_nomix.json 2> gpurun_out/e_bench_nomix.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['config'].get('sort'), d['config'].get('fft'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
