"""Per-entry-point device times of the slab step, P ranks in one process on one GPU (LocalComm):
every pm_slab_* call of rank 0 is bracketed by CUDA events.  Remote = local memory here, so these
are the compute costs of the kernels without NVLink."""
import sys, types, json, collections
import torch
sys.path.insert(0, ".")
import cosmological_particle_mesh_simulation_b200 as pm
import bench

def run(n_parts, n_cells, P, C, transport, steps=4):
    cfg = bench.cfg_namespace(n_parts, n_cells)
    pm.set_config(cfg)
    slab = pm.slab
    comm = slab.LocalComm(P)
    ranks = []
    for r in range(P):
        pl, vl, il = bench.make_particles_slab_gpu(n_parts, n_cells, r, P, 0)
        ranks.append(slab.make_rank_from_local(n_cells, pl, vl, il, r, P, device=0))
        del pl, vl, il
    assert slab.setup_peers(ranks, comm)
    sched = pm.loop_scale_factors(cfg)
    mass = (n_cells / n_parts) ** 3
    for i in range(3):
        slab.slab_step(ranks, comm, *sched[i], mass=mass, cfg=cfg, chunks=C, transport=transport)
    rec = collections.defaultdict(list)
    orig = slab.SlabRank._call
    def timed(self, fn, *args):
        if self.rank != 0:
            return orig(self, fn, *args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(self, fn, *args); e1.record()
        rec[fn].append((e0, e1))
    slab.SlabRank._call = timed
    for i in range(steps):
        slab.slab_step(ranks, comm, *sched[3 + i], mass=mass, cfg=cfg, chunks=C, transport=transport)
    torch.cuda.synchronize()
    slab.SlabRank._call = orig
    out = {k: round(sum(a.elapsed_time(b) for a, b in v) / steps, 4) for k, v in rec.items()}
    for r in ranks:
        r.close()
    pm.release_plans()
    torch.cuda.empty_cache()
    return out

if __name__ == "__main__":
    res = {}
    for (npart, nc, P, C, tr) in [(256, 512, 2, 1, "fused"), (256, 512, 2, 2, "fused"), (256, 512, 2, 1, "peer"), (256, 512, 8, 1, "fused"),
                                  (512, 1024, 2, 1, "fused"), (512, 1024, 2, 2, "fused"), (512, 1024, 8, 1, "fused")]:
        key = f"{npart}^3/{nc}^3 P={P} C={C} {tr}"
        res[key] = run(npart, nc, P, C, tr)
        print(key, json.dumps(res[key]), flush=True)
    json.dump(res, open("gpurun_out/t_prof_slab.json", "w"), indent=1)
