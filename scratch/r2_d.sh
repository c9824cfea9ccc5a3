#!/bin/bash
# Round 2, call D (1 GPU): warp-specialised gather (mbarrier + bulk copy), scan-aggregated tile deposit, staged radix scatter
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/d_pytest.log
tail -30 gpurun_out/d_pytest.log | cut -c1-300
for load in ic evolved clustered; do
  timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/d_bench_$load.json 2> gpurun_out/d_bench_$load.err; echo "bench $load rc=$?"
done
PM_GATHER_WS=0 PM_DEPOSIT=rows timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/d_bench_ic_old.json 2> gpurun_out/d_bench_ic_old.err
PM_GATHER_WS=0 PM_DEPOSIT=rows timeout 600 python bench.py --steps 20 --warmup 3 --particles evolved --no-cpu-baseline --no-e2e > gpurun_out/d_bench_evolved_old.json 2> gpurun_out/d_bench_evolved_old.err
python - <<'PY'
import json
for n in ("ic", "evolved", "clustered", "ic_old", "evolved_old"):
    try:
        d=json.loads(open(f"gpurun_out/d_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, d["config"]["sort"], d["config"]["fft"]["sync_errors"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/d_bench_{n}.err").read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/d_launches_ic.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/d_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit_tiles|k_radix_sort|k_gather_ws' -s 6 -c 6 -o gpurun_out/d_prof_ic \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/d_ncu2.log 2>&1; echo "ncu full rc=$?"
