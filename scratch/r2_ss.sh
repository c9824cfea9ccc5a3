#!/bin/bash
# call SS: NumPy drop-in loop with the caller's arrays page-locked in place and pooled pinned result arrays
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "numpy_drop_in" > gpurun_out/ss_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/ss_pytest.log | cut -c1-400
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ss_bench.json 2> gpurun_out/ss_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/ss_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["value"], {k:round(v["ms_per_step"],3) for k,v in d["e2e_dropin"].items()})
PY
tail -2 gpurun_out/ss_bench.err
