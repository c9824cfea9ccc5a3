#!/bin/bash
# Round 2, call H (1 GPU): CUDA-graph replay on by default (tests + bench), gather phi ring depth 6, zc sweep on the evolved load
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/h_pytest.log
tail -12 gpurun_out/h_pytest.log | cut -c1-400
run() { # name, load, env...
  name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/h_bench_${name}_$load.json 2> gpurun_out/h_bench_${name}_$load.err
}
run main ic PM_X=0
run nograph ic PM_GRAPH=0
run r6 ic PM_LIB=scratch/variants/libpmstep_r6.so
run r6c ic PM_LIB=scratch/variants/libpmstep_r6c.so
run main evolved PM_X=0
run r6 evolved PM_LIB=scratch/variants/libpmstep_r6.so
run r6zc8 evolved PM_LIB=scratch/variants/libpmstep_r6.so PM_GATHER_ZC=8
run r6zc16 evolved PM_LIB=scratch/variants/libpmstep_r6.so PM_GATHER_ZC=16
run r6zc16 ic PM_LIB=scratch/variants/libpmstep_r6.so PM_GATHER_ZC=16
timeout 300 python bench.py --steps 50 --warmup 5 --n-parts 64 --n-cells 128 --no-cpu-baseline --no-e2e > gpurun_out/h_bench_c1_graph.json 2> gpurun_out/h_bench_c1_graph.err
PM_GRAPH=0 timeout 300 python bench.py --steps 50 --warmup 5 --n-parts 64 --n-cells 128 --no-cpu-baseline --no-e2e > gpurun_out/h_bench_c1_nograph.json 2> gpurun_out/h_bench_c1_nograph.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/h_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["config"]["sort"]["mode"], d["config"]["fft"]["sync_errors"], d.get("graph",{}).get("replays_in_timed_region"))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
