#!/bin/bash
# Round 2, call AA (1 GPU): slab row passes through the two-stage kernels: slab tests (P = 1, 2, 4 on one GPU, IPC), entry-point times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/aa_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/aa_pytest.log | cut -c1-400
timeout 900 python - > gpurun_out/aa_prof.log 2>&1 <<'PY'
import sys, json
sys.path.insert(0, "scratch"); sys.path.insert(0, ".")
import prof_slab
res = {}
for (npart, nc, P, C, tr) in [(256, 512, 2, 1, "fused"), (256, 512, 8, 1, "fused"), (512, 1024, 2, 1, "fused"), (512, 1024, 8, 1, "fused")]:
    key = f"{npart}^3/{nc}^3 P={P} C={C} {tr}"
    res[key] = prof_slab.run(npart, nc, P, C, tr)
    print(key, json.dumps(res[key]), flush=True)
json.dump(res, open("gpurun_out/aa_prof_slab.json", "w"), indent=1)
PY
echo "prof rc=$?"; grep "P=" gpurun_out/aa_prof.log | cut -c1-700
