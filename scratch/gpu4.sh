#!/bin/bash
mkdir -p gpurun_out
PM_FFT_FUSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fft2" -s 5 -c 5 -f -o gpurun_out/d_prof_fft2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/d_ncu1.log 2>&1
tail -3 gpurun_out/d_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fft_plane" -s 2 -c 2 -f -o gpurun_out/d_prof_plane python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/d_ncu2.log 2>&1
tail -3 gpurun_out/d_ncu2.log
ls -la gpurun_out/*.ncu-rep
