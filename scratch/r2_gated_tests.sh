#!/bin/bash
# Round-2, first thing (1 GPU, ~2 min): the tests written after round 1's GPU budget was spent.
# gpurun --timeout 600 -- 'bash scratch/r2_gated_tests.sh'
mkdir -p gpurun_out
PM_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_slab_gpu.py -m gpu -q -k experimental -p no:cacheprovider > gpurun_out/z_experimental.log 2>&1
echo "experimental rc=$?"; tail -5 gpurun_out/z_experimental.log
PM_TEST_FULLSIZE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k full_size_parity -p no:cacheprovider > gpurun_out/z_fullsize.log 2>&1
echo "fullsize rc=$?"; tail -5 gpurun_out/z_fullsize.log
