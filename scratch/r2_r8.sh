#!/bin/bash
# Round 2, call R (8 GPUs): bench --gpus 8 as the driver launches it (strong scaling 256^3/512^3 + the weak_scaling sub-record:
# configs[3] on 8 GPUs against configs[2] on one), transports at 2048^3, configs[2] on 8 GPUs, reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r8_topo.txt 2>&1
port=29700
tr() { name=$1; shift; port=$((port+1))
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" > gpurun_out/r8_$name.json 2> gpurun_out/r8_$name.err
  echo "$name rc=$?"
}
tr default --steps 20 --warmup 3
tr c4_fused2_c2 --steps 3 --warmup 2 --no-e2e --no-weak-scaling --n-parts 1024 --n-cells 2048 --transport fused2 --chunks 2
tr c4_peer_c4 --steps 3 --warmup 2 --no-e2e --no-weak-scaling --n-parts 1024 --n-cells 2048 --transport peer --chunks 4
tr c4_fused_c2 --steps 3 --warmup 2 --no-e2e --no-weak-scaling --n-parts 1024 --n-cells 2048 --transport fused --chunks 2
tr c3_fused --steps 5 --warmup 3 --no-e2e --no-weak-scaling --n-parts 512 --n-cells 1024
tr ref --impl reference --steps 1 --warmup 0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d.get('ms_per_step',0),3), d.get('config',{}).get('fft_transport'), d.get('config',{}).get('chunks'), {k:round(v,3) for k,v in d.get('phases_ms_rank0',{}).items()}, (d.get('e2e') or {}).get('value'), d.get('cpu_baseline'))
        if d.get('weak_scaling'): print('   weak:', json.dumps(d['weak_scaling'])[:1500])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
