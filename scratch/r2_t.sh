#!/bin/bash
# Round 2, call T (1 GPU): deposit specialised on the mesh width; final-state ncu launch list + full capture of the top kernels (IC and evolved loads)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/t_pytest.log | cut -c1-400
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/t_bench_${name}_$load.json 2> gpurun_out/t_bench_${name}_$load.err
}
run main ic PM_X=0
run main evolved PM_X=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/t_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["config"]["sort"]["mode"], d["config"].get("gather_items"))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
PM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/t_launches_ic.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t_ncu1.log 2>&1; echo "ncu list ic rc=$?"
PM_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 23000 -c 200 --csv --log-file gpurun_out/t_launches_evolved.csv python bench.py --steps 3 --warmup 3 --particles evolved --no-cpu-baseline --no-e2e > gpurun_out/t_ncu1e.log 2>&1; echo "ncu list evolved rc=$?"
PM_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit_tiles|k_gather_ws|k_radix_sort|k_mover_partition|k_merge_tiles|k_fft_cols' -s 12 -c 8 -o gpurun_out/t_prof_ic \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t_ncu2.log 2>&1; echo "ncu full rc=$?"
