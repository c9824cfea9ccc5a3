#!/bin/bash
# Round 2, call C (1 GPU): own radix sort (sync-free step) + optimised tile deposit: parity suite, bench on the three loads, ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/c_pytest.log
tail -30 gpurun_out/c_pytest.log | cut -c1-300
for load in ic evolved clustered; do
  timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/c_bench_$load.json 2> gpurun_out/c_bench_$load.err; echo "bench $load rc=$?"
done
PM_DEPOSIT=rows timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c_bench_ic_rows.json 2> gpurun_out/c_bench_ic_rows.err
python - <<'PY'
import json
for n in ("ic", "evolved", "clustered", "ic_rows"):
    try:
        d=json.loads(open(f"gpurun_out/c_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, d["config"]["sort"], d["config"]["gather_blocks"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/c_bench_{n}.err").read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c_launches_ic.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit_tiles|k_radix_sort|k_merge_tiles|k_mover_partition' -s 8 -c 8 -o gpurun_out/c_prof_ic \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c_ncu2.log 2>&1; echo "ncu full rc=$?"
