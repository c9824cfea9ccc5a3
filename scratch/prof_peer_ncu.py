"""Two fused slab steps, P = 2 ranks in one process on one GPU (256^3 particles on 512^3 cells), for an
ncu capture of the peer y-pass kernels (k_fft_cols_peer).  "Remote" memory is local here: the capture
shows the kernels' compute / HBM side, not NVLink."""
import sys
sys.path.insert(0, ".")
import torch
import cosmological_particle_mesh_simulation_b200 as pm
import bench

n_parts, n_cells, P = 256, 512, 2
cfg = bench.cfg_namespace(n_parts, n_cells)
pm.set_config(cfg)
slab = pm.slab
comm = slab.LocalComm(P)
ranks = []
for r in range(P):
    pl, vl, il = bench.make_particles_slab_gpu(n_parts, n_cells, r, P, 0)
    ranks.append(slab.make_rank_from_local(n_cells, pl, vl, il, r, P, device=0))
assert slab.setup_peers(ranks, comm)
sched = pm.loop_scale_factors(cfg)
for i in range(2):
    slab.slab_step(ranks, comm, *sched[i], mass=8.0, cfg=cfg, chunks=1, transport="fused")
torch.cuda.synchronize()
print("ok", [r.count for r in ranks], flush=True)
