#!/bin/bash
# Round 2, final single-GPU record after calls GG-NN: smoke, full GPU suite, default bench (as the driver runs it), reference arm, launch list
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/f3_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/f3_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/f3_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/f3_bench_default.json 2> gpurun_out/f3_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/f3_bench_reference.json 2> gpurun_out/f3_bench_reference.err; echo "ref rc=$?"
PM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f3_launches_ic.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f3_ncu1.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f3_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[21:-5], round(d["ms_per_step"],4), d.get("value"), {k:round(v,3) for k,v in (d.get("stages_ms") or {}).items() if v>0.01}, (d.get("e2e") or {}).get("value"), {k:round(v["ms_per_step"],3) for k,v in (d.get("e2e_dropin") or {}).items()}, d.get("cpu_baseline"), d.get("roofline",{}).get("frac"), (d.get("roofline_step") or {}).get("frac"), d.get("gpu_launches"), d.get("clocks"))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
