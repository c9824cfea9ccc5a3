#!/bin/bash
# call GG: slab-by-slab initial conditions -- new tests, the IC tests (kernel signatures changed), timing
mkdir -p gpurun_out
python -m pytest tests/test_slab_ic.py tests/test_gpu_ic.py -x -q -m gpu > gpurun_out/gg_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/gg_pytest.log
timeout 300 python scratch/time_slab_ic.py 256 1 2 4 8 > gpurun_out/gg_slab_ic_256.json 2> gpurun_out/gg_slab_ic_256.err; echo "time256 rc=$?"; cat gpurun_out/gg_slab_ic_256.json
timeout 300 python scratch/time_slab_ic.py 512 8 > gpurun_out/gg_slab_ic_512.json 2> gpurun_out/gg_slab_ic_512.err; echo "time512 rc=$?"; cat gpurun_out/gg_slab_ic_512.json; tail -3 gpurun_out/gg_slab_ic_512.err
