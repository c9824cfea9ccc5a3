#!/bin/bash
# call NN: host-buffer step sorts by mesh row only (two radix passes instead of three): parity tests, timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_step or composed_calls or sort" > gpurun_out/nn_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/nn_pytest.log
PM_HOST_TIMING=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/nn_bench.json 2> gpurun_out/nn_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/nn_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"])
PY
grep "timeline" gpurun_out/nn_bench.err | tail -4
python scratch/pcie_probe.py 2>/dev/null | tail -4
