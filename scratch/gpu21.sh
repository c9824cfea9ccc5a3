#!/bin/bash
# four GPUs: fused transport at configs 2 and 3 (+ one NCCL line for comparison)
mkdir -p gpurun_out
port=29800
run() { name=$1; shift; port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 4 --no-cpu-baseline "$@" > gpurun_out/u_$name.json 2> gpurun_out/u_$name.err
}
run c2_g4_fused --steps 20 --warmup 3
run c2_g4_nccl --steps 20 --warmup 3 --transport nccl
run c3_g4_fused --steps 8 --warmup 3 --n-parts 512 --n-cells 1024
run c3_g4_fused_c2 --steps 8 --warmup 3 --n-parts 512 --n-cells 1024 --chunks 2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/u_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), d['config'].get('fft_transport'), {k:round(v,3) for k,v in d['phases_ms_rank0'].items()}, 'e2e', '%.3g'%d['e2e']['value'], d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
