#!/bin/bash
mkdir -p gpurun_out
timeout 55 ncu --set full --clock-control none --import-source on -k regex:k_fft_cols_peer -s 4 -c 4 -f -o gpurun_out/w_prof_peer python scratch/prof_peer_ncu.py > gpurun_out/w_ncu.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/w_ncu.log; ls -la gpurun_out/w_prof_peer.ncu-rep
