#!/bin/bash
# Round 2, call B (1 GPU): tile deposit + mean-subtracted transform: parity suite, bench on the three loads, A/B against the row deposit
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/b_pytest.log
tail -30 gpurun_out/b_pytest.log
for load in ic evolved clustered; do
  timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/b_bench_$load.json 2> gpurun_out/b_bench_$load.err; echo "bench $load rc=$?"
done
PM_DEPOSIT=rows timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_bench_ic_rows.json 2> gpurun_out/b_bench_ic_rows.err
python - <<'PY'
import json
for n in ("ic", "evolved", "clustered", "ic_rows"):
    try:
        d=json.loads(open(f"gpurun_out/b_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, d["config"]["sort"], d["config"]["gather_blocks"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/b_bench_{n}.err").read()[-1500:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit_tiles|k_deposit_items|k_deposit_slots' -s 3010 -c 6 -o gpurun_out/b_prof_deposit \
  python bench.py --steps 3 --warmup 3 --particles evolved --evolve-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1; echo "ncu rc=$?"
