#!/bin/bash
mkdir -p gpurun_out
PM_LIB=$PWD/scratch/variants/libpmstep_n.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gather_tiled" -s 4 -c 1 -f -o gpurun_out/j_prof_gt python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/j_ncu.log 2>&1
tail -3 gpurun_out/j_ncu.log
ls -la gpurun_out/j_prof_gt.ncu-rep
