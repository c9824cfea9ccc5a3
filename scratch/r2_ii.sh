#!/bin/bash
# call II: slab driver test (+ the driver tests it extends)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_driver.py -x -q -m gpu > gpurun_out/ii_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/ii_pytest.log
