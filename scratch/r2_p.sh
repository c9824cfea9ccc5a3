#!/bin/bash
# Round 2, call P (1 GPU): Poisson options (deconvolution, spectral gradient), gather work list test fix; full suite; default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/p_pytest.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/p_bench_default.json 2> gpurun_out/p_bench_default.err
echo "bench default rc=$?"; tail -3 gpurun_out/p_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/p_bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:round(v,3) for k,v in d["stages_ms"].items() if v>0.01}, d["e2e"]["value"], {k:round(v["ms_per_step"],3) for k,v in d["e2e_dropin"].items()}, d["cpu_baseline"])
PY
