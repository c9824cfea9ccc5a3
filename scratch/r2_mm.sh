#!/bin/bash
# call MM: radix sort histogram phase with all loads of a tile issued first (branch-free): sort tests, stage times, host timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sort or host_step or graph" > gpurun_out/mm_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/mm_pytest.log
PM_HOST_TIMING=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mm_bench.json 2> gpurun_out/mm_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/mm_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("stages_ms"))
PY
grep "timeline" gpurun_out/mm_bench.err | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --particles clustered > gpurun_out/mm_bench_clustered.json 2> gpurun_out/mm_bench_clustered.err; echo "bench clustered rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/mm_bench_clustered.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d.get("stages_ms"))
PY
