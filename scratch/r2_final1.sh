#!/bin/bash
# Round 2, final single-GPU record: smoke, full GPU suite, default bench (as the driver runs it), the three loads, launch list
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f1_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/f1_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/f1_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/f1_bench_default.json 2> gpurun_out/f1_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/f1_bench_reference.json 2> gpurun_out/f1_bench_reference.err; echo "ref rc=$?"
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/f1_bench_${name}_$load.json 2> gpurun_out/f1_bench_${name}_$load.err
}
run main evolved PM_X=0
run main clustered PM_X=0
timeout 300 python bench.py --steps 50 --warmup 5 --n-parts 64 --n-cells 128 --no-cpu-baseline > gpurun_out/f1_bench_c1.json 2> gpurun_out/f1_bench_c1.err
timeout 600 python bench.py --steps 5 --warmup 3 --n-parts 512 --n-cells 1024 --no-cpu-baseline --no-e2e > gpurun_out/f1_bench_c3.json 2> gpurun_out/f1_bench_c3.err
PM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/f1_launches_ic.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f1_ncu1.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f1_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[21:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in (d.get("stages_ms") or {}).items() if v>0.01}, (d.get("e2e") or {}).get("value"), {k:round(v["ms_per_step"],3) for k,v in (d.get("e2e_dropin") or {}).items()}, d.get("cpu_baseline"), d.get("roofline",{}).get("frac"), d.get("roofline_step",{}).get("frac"))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
