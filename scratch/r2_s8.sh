#!/bin/bash
# Round 2, call S (8 GPUs): where the 27 ms of the distributed FFT at 2048^3 go -- per-entry-point device times of rank 0
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 scratch/prof_slab_dist.py 1024 2048 fused 1 gpurun_out/s8_calls_c4_fused_c1.json > gpurun_out/s8_a.log 2>&1; echo rc=$?; tail -2 gpurun_out/s8_a.log | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29812 scratch/prof_slab_dist.py 1024 2048 peer 4 gpurun_out/s8_calls_c4_peer_c4.json > gpurun_out/s8_b.log 2>&1; echo rc=$?; tail -2 gpurun_out/s8_b.log | cut -c1-1500
