#!/bin/bash
# Round 2, call Q (1 GPU): gather CTAs loop over the work list (one CTA per base chunk, further items drawn from a counter)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/q2_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/q2_pytest.log | cut -c1-400
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/q2_bench_${name}_$load.json 2> gpurun_out/q2_bench_${name}_$load.err
}
run items ic PM_X=0
run noitems ic PM_GATHER_ITEMS=0
run items evolved PM_X=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/q2_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[21:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["config"]["sort"]["mode"], d["config"].get("gather_items"), d["config"]["fft"]["sync_errors"])
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
