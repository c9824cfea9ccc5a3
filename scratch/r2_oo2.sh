#!/bin/bash
# call OO (2 GPUs): the NCCL / peer slab tests and the 2-GPU bench line with the final library of the round
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_nccl.py -x -q -m gpu > gpurun_out/oo_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/oo_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/oo_bench_g2.json 2> gpurun_out/oo_bench_g2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/oo_bench_g2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d.get("e2e"), d.get("phases_ms_rank0"), d.get("peer_flag_timeouts"))
PY
tail -2 gpurun_out/oo_bench_g2.err
