#!/bin/bash
# Round 2, last validation of the final tree (1 GPU): smoke, full GPU suite, default bench
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f2_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/f2_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/f2_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/f2_bench_default.json 2> gpurun_out/f2_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/f2_bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_step"]["frac"], d["gpu_launches"], d["clocks"])
PY
