#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/n_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/n_pytest.log
tail -4 gpurun_out/n_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/n_bench_n1.json 2> gpurun_out/n_bench_n1.err
tail -c 600 gpurun_out/n_bench_n1.json
timeout 300 python bench.py --steps 5 --warmup 3 --n-parts 512 --n-cells 1024 --no-cpu-baseline > gpurun_out/n_bench_c3.json 2> gpurun_out/n_bench_c3.err
timeout 300 python bench.py --steps 50 --warmup 5 --n-parts 64 --n-cells 128 --no-cpu-baseline > gpurun_out/n_bench_c1.json 2> gpurun_out/n_bench_c1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/n_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_deposit|k_fft|k_gather|k_merge|k_mover_part" -s 27 -c 9 -f -o gpurun_out/n_prof_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_ncu_full.log 2>&1
ls -la gpurun_out/n_prof_full.ncu-rep
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/n_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()}, d.get('cpu_baseline'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
