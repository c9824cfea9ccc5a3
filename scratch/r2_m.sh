#!/bin/bash
# Round 2, call M (1 GPU): spike A/B test, bit-exact Green table, host-step timeline, gather planes-per-CTA sweep on the evolved load
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/m_pytest.log | cut -c1-400
PM_HOST_TIMING=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/m_bench_default.json 2> gpurun_out/m_bench_default.err
echo "bench default rc=$?"; grep timeline gpurun_out/m_bench_default.err | tail -4
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/m_bench_${name}_$load.json 2> gpurun_out/m_bench_${name}_$load.err
}
run main evolved PM_X=0
run zc16 evolved PM_GATHER_ZC=16
run zc8 evolved PM_GATHER_ZC=8
run zc4 evolved PM_GATHER_ZC=4
run zc16 ic PM_GATHER_ZC=16
run tiled evolved PM_GATHER_WS=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["config"]["sort"]["mode"], (d.get("e2e") or {}).get("value"), {k:round(v["ms_per_step"],3) for k,v in (d.get("e2e_dropin") or {}).items()})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
