#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ic.py -m gpu -q -x > gpurun_out/o_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/o_pytest.log
tail -40 gpurun_out/o_pytest.log
