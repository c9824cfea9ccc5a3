"""Per-entry-point device times of the slab step in a REAL multi-GPU run (torchrun, one rank per GPU):
every pm_slab_* call of rank 0 is bracketed by CUDA events on the stream it is issued on.
    torchrun --nproc-per-node 8 scratch/prof_slab_dist.py N_PARTS N_CELLS TRANSPORT CHUNKS OUT.json"""
import sys, os, json, collections
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmological_particle_mesh_simulation_b200 as pm
import bench

n_parts, n_cells, transport, C, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), sys.argv[5]
local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
cfg = bench.cfg_namespace(n_parts, n_cells)
pm.set_config(cfg)
slab = pm.slab
comm = slab.DistComm()
pl, vl, il = bench.make_particles_slab_gpu(n_parts, n_cells, rank, world, local)
ranks = [slab.make_rank_from_local(n_cells, pl, vl, il, rank, world, device=local, total_particles=n_parts ** 3)]
del pl, vl, il
torch.cuda.empty_cache()
assert slab.setup_peers(ranks, comm)
aux = slab.setup_ghost_peers(ranks, comm)
kw = dict(mass=8.0, cfg=cfg, chunks=C or None, transport=transport, ghosts="peer" if aux else "nccl", migrate="peer" if aux else "nccl")
sched = pm.loop_scale_factors(cfg)
for i in range(2):
    slab.slab_step(ranks, comm, *sched[i], **kw)
dist.barrier(); torch.cuda.synchronize()
rec = collections.defaultdict(list)
orig = slab.SlabRank._call
def timed(self, fn, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(self, fn, *args); e1.record()
    rec[fn].append((e0, e1))
    return r
slab.SlabRank._call = timed
steps = 3
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(steps):
    slab.slab_step(ranks, comm, *sched[2 + i], **kw)
t1.record()
torch.cuda.synchronize()
slab.SlabRank._call = orig
res = {k: round(sum(a.elapsed_time(b) for a, b in v) / steps, 4) for k, v in rec.items()}
res["_step_ms"] = round(t0.elapsed_time(t1) / steps, 4)
res["_calls_per_step"] = {k: len(v) // steps for k, v in rec.items()}
allres = [None] * world
dist.all_gather_object(allres, res)
if rank == 0:
    json.dump({"config": f"{n_parts}^3/{n_cells}^3 P={world} C={C} {transport}", "rank0": allres[0], "max_over_ranks": {k: max(r[k] for r in allres) for k in res if not k.startswith("_calls")}}, open(out, "w"), indent=1)
    print(json.dumps(allres[0]))
slab.release_peers(ranks, comm)
for r in ranks: r.close()
dist.barrier(); dist.destroy_process_group()
