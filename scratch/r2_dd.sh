#!/bin/bash
# Round 2, call DD (1 GPU): copy-engine transposes of the "peer" transport: slab tests with PM_PEER_DMA=1 (P ranks in one process)
mkdir -p gpurun_out
PM_PEER_DMA=1 timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/dd_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/dd_pytest.log | cut -c1-600
