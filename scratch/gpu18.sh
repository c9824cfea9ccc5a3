#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/r_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r_pytest.log; tail -25 gpurun_out/r_pytest.log
