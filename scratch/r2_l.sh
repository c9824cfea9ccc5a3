#!/bin/bash
# Round 2, call L (1 GPU): reference sin^2 table + exact reciprocal in G, sector-wide un-permute, chunked host download
mkdir -p gpurun_out
python scratch/pcie_probe.py > gpurun_out/l_pcie.txt 2>&1; cat gpurun_out/l_pcie.txt
for be in own cufft f64; do PM_FFT_BACKEND=$be timeout 300 python scratch/diag_spike.py; done > gpurun_out/l_diag_spike.txt 2>&1
cut -c1-420 gpurun_out/l_diag_spike.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/l_pytest.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/l_bench_default.json 2> gpurun_out/l_bench_default.err
echo "bench default rc=$?"; tail -3 gpurun_out/l_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/l_bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["e2e"]["value"], {k:round(v["ms_per_step"],3) for k,v in d["e2e_dropin"].items()})
PY
