#!/bin/bash
# Round 2, call K (1 GPU): float64 diagnostic transforms on the spike fixtures, lazy drop-in, full suite, default bench
mkdir -p gpurun_out
for be in own cufft f64; do PM_FFT_BACKEND=$be timeout 300 python scratch/diag_spike.py; done > gpurun_out/k_diag_spike.txt 2>&1
cat gpurun_out/k_diag_spike.txt | cut -c1-700
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/k_pytest.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/k_bench_default.json 2> gpurun_out/k_bench_default.err
echo "bench default rc=$?"; tail -3 gpurun_out/k_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/k_bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"], json.dumps(d["e2e_dropin"])[:1500])
PY
