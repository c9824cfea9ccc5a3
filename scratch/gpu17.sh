#!/bin/bash
# two GPUs: NCCL + peer-memory slab test, bench with both transports
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_nccl.py -m gpu -q -p no:cacheprovider > gpurun_out/q_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/q_pytest.log; tail -5 gpurun_out/q_pytest.log
for tr in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 2 --steps 20 --warmup 3 --transport $tr > gpurun_out/q_bench_g2_$tr.json 2> gpurun_out/q_bench_g2_$tr.err
echo "bench $tr rc=$?"; tail -3 gpurun_out/q_bench_g2_$tr.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 --n-parts 512 --n-cells 1024 --no-cpu-baseline > gpurun_out/q_bench_c3_g2_auto.json 2> gpurun_out/q_bench_c3_g2_auto.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 10 --warmup 3 --n-parts 512 --n-cells 1024 --no-cpu-baseline --transport nccl > gpurun_out/q_bench_c3_g2_nccl.json 2> gpurun_out/q_bench_c3_g2_nccl.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/q_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), d['config'].get('fft_transport'), {k:round(v,3) for k,v in d['phases_ms_rank0'].items()}, 'e2e', '%.3g'%d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
