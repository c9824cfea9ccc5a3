#!/bin/bash
# call QQ: pm_step_host waits for the caller's work on the device at entry; full GPU suite, smoke, bench with timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --maxfail=5 -p no:cacheprovider > gpurun_out/qq_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/qq_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
PM_HOST_TIMING=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/qq_bench.json 2> gpurun_out/qq_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/qq_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("roofline_step",{}).get("frac"))
PY
grep "timeline" gpurun_out/qq_bench.err | tail -3
