#!/bin/bash
# Round 2, call F (1 GPU): deposit 384 threads + load pipeline, radix column sums vectorised, gather YB=2 variants, drop-in session path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/f_pytest.log
tail -12 gpurun_out/f_pytest.log | cut -c1-300
run() { # name, env...
  name=$1; shift
  for load in ic evolved; do
    env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/f_bench_${name}_$load.json 2> gpurun_out/f_bench_${name}_$load.err
  done
}
run main PM_X=0
run yb2 PM_LIB=scratch/variants/libpmstep_yb2.so
run yb2b PM_LIB=scratch/variants/libpmstep_yb2b.so
timeout 600 python bench.py --steps 20 --warmup 3 --particles clustered --no-cpu-baseline --no-e2e > gpurun_out/f_bench_main_clustered.json 2> gpurun_out/f_bench_main_clustered.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["config"]["sort"]["mode"], d["config"]["fft"]["sync_errors"])
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit_tiles|k_radix_sort' -s 4 -c 4 -o gpurun_out/f_prof_ic \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu2.log 2>&1; echo "ncu full rc=$?"
