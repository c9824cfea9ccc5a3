#!/bin/bash
# Round 2, call X (1 GPU): deposit work-item thresholds on the z = 0 snapshot; slot conversion with loads in flight
mkdir -p gpurun_out
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/x_bench_${name}_$load.json 2> gpurun_out/x_bench_${name}_$load.err
}
run h8192_i8192 evolved PM_X=0
run h4096_i4096 evolved PM_DEP_HEAVY=4096 PM_DEP_ITEM=4096
run h4096_i2048 evolved PM_DEP_HEAVY=4096 PM_DEP_ITEM=2048
run h16384_i8192 evolved PM_DEP_HEAVY=16384 PM_DEP_ITEM=8192
run h3072_i4096 evolved PM_DEP_HEAVY=3072 PM_DEP_ITEM=4096
run h8192_i8192 ic PM_X=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/x_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
