#!/bin/bash
# call PP: host-buffer step uploads y and z first; row keys and the row sort run under the upload of x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_step or composed_calls or numpy_drop_in" > gpurun_out/pp_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pp_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
PM_HOST_TIMING=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/pp_bench.json 2> gpurun_out/pp_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/pp_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"])
PY
grep "timeline" gpurun_out/pp_bench.err | tail -4
