#!/bin/bash
# call HH (2 GPUs): slab initial conditions over NCCL, then the slab bench started from them
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_nccl.py -x -q -m gpu > gpurun_out/hh_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/hh_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 --particles zeldovich --no-e2e > gpurun_out/hh_bench_g2_zeldovich.json 2> gpurun_out/hh_bench_g2_zeldovich.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/hh_bench_g2_zeldovich.json; tail -3 gpurun_out/hh_bench_g2_zeldovich.err
