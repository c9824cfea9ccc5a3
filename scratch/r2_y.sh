#!/bin/bash
# Round 2, call Y (1 GPU): deposit work-item thresholds, upper range, on the z = 0 snapshot and the blob
mkdir -p gpurun_out
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/y_bench_${name}_$load.json 2> gpurun_out/y_bench_${name}_$load.err
}
run h16384_i8192 evolved PM_DEP_HEAVY=16384 PM_DEP_ITEM=8192
run h32768_i8192 evolved PM_DEP_HEAVY=32768 PM_DEP_ITEM=8192
run h32768_i16384 evolved PM_DEP_HEAVY=32768 PM_DEP_ITEM=16384
run h65536_i16384 evolved PM_DEP_HEAVY=65536 PM_DEP_ITEM=16384
run h24576_i8192 evolved PM_DEP_HEAVY=24576 PM_DEP_ITEM=8192
run nosplit evolved PM_DEP_HEAVY=1000000000 PM_DEP_ITEM=8192
run h16384_i8192 clustered PM_DEP_HEAVY=16384 PM_DEP_ITEM=8192
run h32768_i8192 clustered PM_DEP_HEAVY=32768 PM_DEP_ITEM=8192
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/y_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
