#!/bin/bash
# Round 2, call Z (1 GPU): deposit work items as slices of a tile's concatenated runs; thresholds again; gather threshold factor
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider -k "density or resident or incremental or full_size or tiled or graph or slab" > gpurun_out/z_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/z_pytest.log | cut -c1-400
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/z_bench_${name}_$load.json 2> gpurun_out/z_bench_${name}_$load.err
}
run h16384_i8192 evolved PM_X=0
run h8192_i8192 evolved PM_DEP_HEAVY=8192 PM_DEP_ITEM=8192
run h8192_i4096 evolved PM_DEP_HEAVY=8192 PM_DEP_ITEM=4096
run h16384_i4096 evolved PM_DEP_HEAVY=16384 PM_DEP_ITEM=4096
run h12288_i6144 evolved PM_DEP_HEAVY=12288 PM_DEP_ITEM=6144
run gt15 evolved PM_GATHER_T_FACTOR=1.5
run gt3 evolved PM_GATHER_T_FACTOR=3
run gt4 evolved PM_GATHER_T_FACTOR=4
run gt4 ic PM_GATHER_T_FACTOR=4
run main ic PM_X=0
run main clustered PM_X=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/z_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["config"].get("gather_items",{}).get("heavy"))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
