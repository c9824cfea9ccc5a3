#!/bin/bash
# Round 2, call N (1 GPU): pm_step_host with serialised uploads + physical reorder; parity of the host path; e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider -k "host or drop_in or fused_step or numpy or driver or errors" > gpurun_out/n_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/n_pytest.log | cut -c1-400
PM_HOST_TIMING=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/n_bench_default.json 2> gpurun_out/n_bench_default.err
echo "bench default rc=$?"; grep timeline gpurun_out/n_bench_default.err | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/n_bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["value"], {k:round(v["ms_per_step"],3) for k,v in d["e2e_dropin"].items()})
PY
