#!/bin/bash
# call RR: the default bench line (as the driver runs it) with the library as committed at the end of the round
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/rr_bench_default.json 2> gpurun_out/rr_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/rr_bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["cpu_baseline"], d["roofline"]["frac"], d["roofline_step"]["frac"], d["gpu_launches"], d["clocks"])
PY
