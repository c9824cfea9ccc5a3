#!/bin/bash
# Round 2, call BB (1 GPU): gather with two particles per consumer lane (PP = 2), 4 / 6 / 8 consumer warps
mkdir -p gpurun_out
for v in pp2cw4 pp2cw6 pp2cw8; do
  PM_LIB=scratch/variants/libpmstep_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "tiled_gather or graph_replay or full_size_parity" > gpurun_out/bb_pytest_$v.log 2>&1
  echo "$v pytest rc=$?"; tail -2 gpurun_out/bb_pytest_$v.log | cut -c1-300
done
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/bb_bench_${name}_$load.json 2> gpurun_out/bb_bench_${name}_$load.err
}
run main ic PM_X=0
for v in pp2cw4 pp2cw6 pp2cw8; do run $v ic PM_LIB=scratch/variants/libpmstep_$v.so; done
run main evolved PM_X=0
for v in pp2cw4 pp2cw6 pp2cw8; do run $v evolved PM_LIB=scratch/variants/libpmstep_$v.so; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bb_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[21:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
