#!/bin/bash
# eight GPUs, one shot: 1024^3/2048^3 and 256^3/512^3 with the fused peer-memory transposes
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29901 bench.py --gpus 8 --no-cpu-baseline --no-e2e --steps 5 --warmup 3 --n-parts 1024 --n-cells 2048 > gpurun_out/v_c4_g8_fused.json 2> gpurun_out/v_c4_g8_fused.err
echo "c4 rc=$?"
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29902 bench.py --gpus 8 --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/v_c2_g8_fused.json 2> gpurun_out/v_c2_g8_fused.err
echo "c2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/v_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],3), d['config'].get('fft_transport'), {k:round(v,3) for k,v in d['phases_ms_rank0'].items()}, 'e2e', d['e2e'] and '%.3g'%d['e2e']['value'], d['gpu_launches'], d['roofline_step'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
