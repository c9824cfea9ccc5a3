#!/bin/bash
# validation of the driver/snapshot rows and the peer-memory transposes on one GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/p_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/p_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/p_pytest.log
tail -30 gpurun_out/p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/p_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/p_bench_n1.json 2> gpurun_out/p_bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/p_bench_n1.json
