#!/bin/bash
# Round 2, call A (1 GPU): full GPU suite incl. the un-gated full-size parity tests and the experimental
# slab test; bench lines for the IC-like, clustered and evolved loads; ncu launch list + full capture
# of deposit/gather on the evolved load.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt
free -g | head -2 >> gpurun_out/a_gpu.txt; nproc >> gpurun_out/a_gpu.txt
PM_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --durations=8 > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/a_pytest.log
tail -25 gpurun_out/a_pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench_ic.json 2> gpurun_out/a_bench_ic.err; echo "bench ic rc=$?"
timeout 400 python bench.py --steps 20 --warmup 3 --particles clustered --no-cpu-baseline > gpurun_out/a_bench_clustered.json 2> gpurun_out/a_bench_clustered.err; echo "bench clustered rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --particles evolved --no-cpu-baseline > gpurun_out/a_bench_evolved.json 2> gpurun_out/a_bench_evolved.err; echo "bench evolved rc=$?"
python - <<'PY'
import json
for n in ("ic", "clustered", "evolved"):
    try:
        d=json.loads(open(f"gpurun_out/a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, d["config"]["sort"], d["config"]["gather_blocks"], "e2e", (d["e2e"] or {}).get("value"))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/a_bench_{n}.err").read()[-1500:])
PY
# launch list on the clustered load (the evolved one needs 999 steps under ncu: too slow)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/a_launches_clustered.csv \
  python bench.py --steps 3 --warmup 3 --particles clustered --no-cpu-baseline --no-e2e > gpurun_out/a_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit_rows|k_gather_tiled' -s 8 -c 4 -o gpurun_out/a_prof_clustered \
  python bench.py --steps 3 --warmup 3 --particles clustered --no-cpu-baseline --no-e2e > gpurun_out/a_ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
