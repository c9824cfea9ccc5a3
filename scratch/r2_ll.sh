#!/bin/bash
# call LL: stencil sums as one float4 record per particle (host-buffer step); ncu of the cold-start radix sort
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_step or composed_calls" > gpurun_out/ll_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/ll_pytest.log
PM_HOST_TIMING=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ll_bench.json 2> gpurun_out/ll_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/ll_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"])
PY
grep "timeline" gpurun_out/ll_bench.err | tail -3
PM_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_radix_sort' -c 2 -o gpurun_out/ll_prof_radix \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ll_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ll_prof_radix.ncu-rep
