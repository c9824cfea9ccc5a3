#!/bin/bash
# Round 2, call V (1 GPU): smoke() with the resident part; fused-z-pass kernel choice at 1024^3; launch list of the evolved load; merge kernel under ncu
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/v_smoke.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --n-parts 512 --n-cells 1024 --no-cpu-baseline --no-e2e > gpurun_out/v_bench_c3_$name.json 2> gpurun_out/v_bench_c3_$name.err
}
run main PM_X=0
run zmix0 PM_FFT_ZMIX=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
PM_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 17600 -c 120 --csv --log-file gpurun_out/v_launches_evolved.csv python bench.py --steps 3 --warmup 3 --particles evolved --no-cpu-baseline --no-e2e > gpurun_out/v_ncu1e.log 2>&1; echo "ncu list evolved rc=$?"
PM_GRAPH=0 timeout 600 ncu --set full --clock-control none -k regex:'k_merge_tiles|k_mover_partition' -s 6 -c 4 -o gpurun_out/v_prof_sort python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v_ncu2.log 2>&1; echo "ncu full rc=$?"
