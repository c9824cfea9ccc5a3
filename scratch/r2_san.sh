#!/bin/bash
# Round 2, sanitizer call (1 GPU): compute-sanitizer memcheck + racecheck over the small-mesh tests of the kernels written this round
mkdir -p gpurun_out
T='tests/test_gpu_parity.py'
K='tiled_gather_equals_flat_gather_bit_for_bit and 128 or incremental_sort_equals_full_sort_bit_for_bit and 32-64 or density_matches_golden or graph_replay and 64-128 or resident_state_matches and free32 or fused_step_and_host_step'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 300 python -m pytest $T -m gpu -q -x -p no:cacheprovider -k "$K" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/san_memcheck.log | tail -8
K2='tiled_gather_equals_flat_gather_bit_for_bit and 128-64-6000 or incremental_sort_equals_full_sort_bit_for_bit and 32-64 or density_matches_golden and clustered32'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 300 python -m pytest $T -m gpu -q -x -p no:cacheprovider -k "$K2" > gpurun_out/san_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/san_racecheck.log | tail -8
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 --launch-timeout 300 python -m pytest $T -m gpu -q -x -p no:cacheprovider -k "$K2" > gpurun_out/san_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_synccheck.log | tail -4
