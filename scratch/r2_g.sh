#!/bin/bash
# Round 2, call G (1 GPU): full suite (slab peer migration + ghosts by default, deposit v4, radix column scan), bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/g_pytest.log
tail -25 gpurun_out/g_pytest.log | cut -c1-400
for load in ic evolved clustered; do
  timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/g_bench_$load.json 2> gpurun_out/g_bench_$load.err
done
PM_GATHER_WS=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/g_bench_tiledgather_ic.json 2> gpurun_out/g_bench_tiledgather_ic.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["config"]["sort"]["mode"], d["config"]["fft"]["sync_errors"])
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/g_launches_ic.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/g_ncu1.log 2>&1; echo "ncu list rc=$?"
