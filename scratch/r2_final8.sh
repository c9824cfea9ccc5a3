#!/bin/bash
# Round 2, final 8-GPU record: bench --gpus 8 as the driver launches it (incl. weak_scaling), configs[2] on 8 GPUs, 4-GPU subset
mkdir -p gpurun_out
port=29900
tr() { n=$1; name=$2; shift 2; port=$((port+1))
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/f8_$name.json 2> gpurun_out/f8_$name.err
  echo "$name rc=$?"
}
tr 8 g8_default --steps 20 --warmup 3
tr 8 g8_c3 --steps 5 --warmup 3 --no-e2e --no-weak-scaling --n-parts 512 --n-cells 1024
tr 4 g4_default --steps 20 --warmup 3
tr 4 g4_c3 --steps 5 --warmup 3 --no-e2e --n-parts 512 --n-cells 1024
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d.get('ms_per_step',0),3), d.get('config',{}).get('fft_transport'), {k:round(v,3) for k,v in d.get('phases_ms_rank0',{}).items()}, (d.get('e2e') or {}).get('value'))
        if d.get('weak_scaling'): w=d['weak_scaling']; print('   weak:', w.get('ms_per_step_1gpu'), w.get('ms_per_step_all_gpus'), w.get('parallel_efficiency'), w.get('phases_ms_rank0'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
