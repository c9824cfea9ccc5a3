#!/bin/bash
# Round 2, call U (2 GPUs): full GPU suite on a two-GPU box (two-device test, NCCL slab test), bitmap merge, peer-wait timeout variable; bench at 1 and 2 GPUs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/u_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/u_pytest.log | cut -c1-400
run() { name=$1; load=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --particles $load --no-cpu-baseline --no-e2e > gpurun_out/u_bench_${name}_$load.json 2> gpurun_out/u_bench_${name}_$load.err
}
run main ic PM_X=0
run main evolved PM_X=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/u_bench_g2.json 2> gpurun_out/u_bench_g2.err; echo "g2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/u_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[20:-5], round(d["ms_per_step"],4), {k:round(v,3) for k,v in (d.get("stages_ms") or d.get("phases_ms_rank0")).items() if v>0.01})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
