#!/bin/bash
# Round-2 experiment (1 GPU, ~1.5 min): the step under the clustered microbench load (BASELINE configs[4]):
# per-stage times show how the deposit's segmented scan, the sort and the gather cope with skew.
# gpurun --timeout 300 -- 'bash scratch/r2_clustered.sh'
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 3 --particles clustered --no-cpu-baseline > gpurun_out/y_clustered.json 2> gpurun_out/y_clustered.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/y_ic.json 2> gpurun_out/y_ic.err
python - <<'PY'
import json
for n in ("clustered", "ic"):
    d=json.loads(open(f"gpurun_out/y_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()})
PY
