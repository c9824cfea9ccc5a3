#!/bin/bash
# Round-2 experiment (8 GPUs, ~2.5 min): which transport wins at 1024^3/2048^3 and at 256^3/512^3.
# gpurun --gpus 8 --timeout 400 -- 'bash scratch/r2_multi8.sh'
mkdir -p gpurun_out
port=29950
run() { name=$1; shift; port=$((port+1))
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus 8 --no-cpu-baseline --no-e2e "$@" > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err
}
run c4_peer_c4   --steps 5 --warmup 3 --n-parts 1024 --n-cells 2048 --transport peer --chunks 4
run c4_fused2_c2 --steps 5 --warmup 3 --n-parts 1024 --n-cells 2048 --transport fused2 --chunks 2
run c2_fused2_c2 --steps 20 --warmup 3 --transport fused2 --chunks 2
run c2_peer_c1   --steps 20 --warmup 3 --transport peer --chunks 1
run c2_fused_gp  --steps 20 --warmup 3 --transport fused --ghosts peer     # after PM_TEST_EXPERIMENTAL=1 pytest passed on one GPU
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/x_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), d['config'].get('fft_transport'), {k:round(v,3) for k,v in d['phases_ms_rank0'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
