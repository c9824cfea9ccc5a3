#!/bin/bash
# Round 2, call J (2 GPUs): two-rank NCCL/peer test on real NVLink, bench --gpus 2 as the driver launches it, weak_scaling code path at small sizes
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/j_topo.txt 2>&1
timeout 600 python -m pytest tests/test_slab_nccl.py tests/test_slab_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/j_pytest.log | cut -c1-600
tr() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 "$@" > gpurun_out/j_$name.json 2> gpurun_out/j_$name.err
  echo "$name rc=$?"
}
tr bench_g2 --steps 20 --warmup 3
tr bench_g2_weak --steps 10 --warmup 3 --no-e2e --weak-scaling-sizes 128,256,256,512
tr bench_g2_c2 --steps 5 --warmup 3 --no-e2e --n-parts 512 --n-cells 1024
tr bench_g2_ref --impl reference --steps 1 --warmup 0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/j_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d.get('ms_per_step',0),3), d.get('config',{}).get('fft_transport'), {k:round(v,3) for k,v in d.get('phases_ms_rank0',{}).items()}, (d.get('e2e') or {}).get('value'), d.get('weak_scaling'), d.get('cpu_baseline'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
