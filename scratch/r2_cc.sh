#!/bin/bash
# Round 2, call CC (1 GPU): two-stage row kernel at 2048 points: equivalence test and device times of one rank's row passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/cc_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/cc_pytest.log | cut -c1-400
for a in "512 2" "1024 8" "2048 8"; do timeout 300 python scratch/rows_time.py $a; done > gpurun_out/cc_rows_time.txt 2>&1
cat gpurun_out/cc_rows_time.txt | grep mesh
