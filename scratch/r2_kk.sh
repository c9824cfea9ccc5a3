#!/bin/bash
# call KK: pm_step_host split (sums-only gather at the original index + push in the caller's order behind the
# chunked velocity upload): parity tests, e2e with and without the split, timelines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_step or composed_calls or full_size" > gpurun_out/kk_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/kk_pytest.log
for v in 1 0; do
  PM_HOST_SPLIT=$v PM_HOST_TIMING=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/kk_bench_split$v.json 2> gpurun_out/kk_bench_split$v.err; echo "bench split=$v rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/kk_bench_split$v.json").read().strip().splitlines()[-1])
print("split=$v", d["ms_per_step"], d["value"], d["e2e"])
PY
  grep "timeline" gpurun_out/kk_bench_split$v.err | tail -3
done
