#!/bin/bash
# Round 2, call EE (8 GPUs): configs[3] with the "peer" transport's transposes on the copy engines (PM_PEER_DMA=1), 4 and 8 chunks
mkdir -p gpurun_out
port=29950
tr() { name=$1; shift; port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --no-e2e --no-weak-scaling --n-parts 1024 --n-cells 2048 --steps 3 --warmup 2 "$@" > gpurun_out/ee8_$name.json 2> gpurun_out/ee8_$name.err
  echo "$name rc=$?"
}
export PM_PEER_DMA=1
tr dma_c4 --transport peer --chunks 4
tr dma_c8 --transport peer --chunks 8
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ee8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d.get('ms_per_step',0),3), d.get('config',{}).get('fft_transport'), d.get('config',{}).get('chunks'), {k:round(v,3) for k,v in d.get('phases_ms_rank0',{}).items()}, d.get('mass_conservation_rel_err'), d.get('peer_flag_timeouts'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
