#!/bin/bash
# Round 2, call W (1 GPU): deposit slots converted by the last finisher (1024 slots), smoke with the resident part; suite; three loads
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/w_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/w_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/w_pytest.log | cut -c1-400
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/w_bench_${name}_$load.json 2> gpurun_out/w_bench_${name}_$load.err
}
run main ic PM_X=0
run main evolved PM_X=0
run main clustered PM_X=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/w_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["config"]["sort"]["mode"])
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
