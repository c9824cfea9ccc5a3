#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "fft or fused or potential or incremental or resident or golden or free" > gpurun_out/f_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/f_pytest.log
tail -15 gpurun_out/f_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
PM_FFT_V3=2 timeout 300 $B > gpurun_out/f_bench_v3.json 2> gpurun_out/f_bench_v3.err
PM_FFT_V3=2 PM_GATHER_COUNT=0 timeout 300 $B > gpurun_out/f_bench_v3_nocount.json 2> gpurun_out/f_bench_v3_nocount.err
PM_FFT_V3=2 PM_FFT_V3Z=256 timeout 300 $B > gpurun_out/f_bench_v3z256.json 2> gpurun_out/f_bench_v3z256.err
PM_FFT_V3=2 PM_FFT_V3Z=512 timeout 300 $B > gpurun_out/f_bench_v3z512.json 2> gpurun_out/f_bench_v3z512.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
