#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_nccl.py -m gpu -q -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/s_pytest.log; tail -5 gpurun_out/s_pytest.log
port=29700
run() { # name, args...
  name=$1; shift; port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --no-cpu-baseline "$@" > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
}
run c2_fused_auto --steps 20 --warmup 3 --transport fused
run c2_fused_c1 --steps 20 --warmup 3 --transport fused --chunks 1
run c2_fused_c2 --steps 20 --warmup 3 --transport fused --chunks 2
run c2_peer_c2 --steps 20 --warmup 3 --transport peer --chunks 2
run c3_fused_auto --steps 8 --warmup 3 --n-parts 512 --n-cells 1024 --transport fused
run c3_fused_c2 --steps 8 --warmup 3 --n-parts 512 --n-cells 1024 --transport fused --chunks 2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), d['config'].get('fft_transport'), {k:round(v,3) for k,v in d['phases_ms_rank0'].items()}, 'e2e', '%.3g'%d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
