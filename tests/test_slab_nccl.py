"""Two-process NCCL run of the slab step against the single-GPU step (needs >= 2 GPUs; skipped on
the one-GPU box, where tests/test_slab_gpu.py covers the same kernels through the rank loop)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import sys, types, torch, torch.distributed as dist
sys.path.insert(0, %r)
import cosmological_particle_mesh_simulation_b200 as pm
from oracle import oracle as O
local = int(__import__("os").environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda:%%d" %% local))
n_parts, n_cells = 64, 128
cfg = types.SimpleNamespace(**O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100).__dict__)
pm.set_config(cfg)
pos_h, vel_h = O.lattice_ic(n_parts, n_cells, seed=11, vel_rms=0.3)
pos, vel = torch.from_numpy(pos_h).cuda(), torch.from_numpy(vel_h).cuda()
comm = pm.slab.DistComm()
ranks = pm.slab.make_ranks(n_cells, pos, vel, comm, device=local)
peer_ok = pm.slab.setup_peers(ranks, comm)     # CUDA IPC over NVLink; the NCCL path needs no set-up
assert peer_ok, "peer-memory transport could not be set up on this box"
assert pm.slab.setup_ghost_peers(ranks, comm), "mapping of the ghost / migration buffers failed"
ref_p, ref_v = pos.clone(), vel.clone()
a, da = 0.3, 0.0099
for s in range(6):
    pm.step(ref_p, ref_v, a, da, mass=8.0)
    pm.slab.slab_step(ranks, comm, a, da, mass=8.0, cfg=cfg, chunks=(1, 2)[s %% 2],
                      transport=("fused", "peer", "nccl")[s %% 3], ghosts=("peer", "nccl")[s %% 2],
                      migrate=("peer", "peer", "nccl", "nccl")[s %% 4])
    a += da
torch.cuda.synchronize()
assert ranks[0].peer_timeouts() == 0
got_p, got_v = pm.slab.collect(ranks, comm, pos.shape[1])
d = torch.remainder(got_p.double() - ref_p.double() + n_cells / 2, n_cells) - n_cells / 2
ep = float(d.norm() / ref_p.double().norm()); ev = float((got_v.double() - ref_v.double()).norm() / ref_v.double().norm())
assert ep <= 1e-6 and ev <= 1e-5, (ep, ev)
tot = torch.tensor([ranks[0].count], device="cuda"); dist.all_reduce(tot)
assert int(tot.item()) == pos.shape[1]
pm.slab.release_peers(ranks, comm)
for r in ranks: r.close()
dist.barrier(); dist.destroy_process_group()
sys.stdout.write("rank" + str(comm.rank) + "-ok %%.2e %%.2e\n" %% (ep, ev)); sys.stdout.flush()
'''


def test_two_rank_nccl_slab_step_matches_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % REPO)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("-ok") == 2, out.stdout


IC_WORKER = r'''
import sys, types, importlib, torch, torch.distributed as dist
sys.path.insert(0, %r)
import cosmological_particle_mesh_simulation_b200 as pm
from oracle import oracle_ic as IC
local = int(__import__("os").environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda:%%d" %% local))
n_parts, n_cells = 64, 128
d = dict(IC.ICConfig(N_PARTS=n_parts, N_CELLS=n_cells).__dict__); d.update(RANDOM_SEED=38, STEPS=100, A_END=1.0)
cfg = types.SimpleNamespace(**d)
pm.set_config(cfg)
comm = pm.slab.DistComm()
# every rank also runs the single-GPU generator: the slab run must hand it that generator's particles
G = importlib.import_module("cosmological_particle_mesh_simulation_b200.gaussian_random_field")
Z = importlib.import_module("cosmological_particle_mesh_simulation_b200.zeldovich")
pos, vel = Z.zeldovich(G.gaussian_random_field())
(p, v, ids), = pm.slab_ic.slab_initial_conditions(comm, device=local)
i = ids.long()
assert bool((pm.slab.slab_of_particles(p[2], n_cells, comm.nranks) == comm.rank).all())
dd = torch.remainder(p.double() - pos[:, i].double() + n_cells / 2, n_cells) - n_cells / 2
ep = float(dd.norm() / pos.double().norm()); ev = float((v.double() - vel[:, i].double()).norm() / vel.double().norm())
assert ep <= 1e-6 and ev <= 1e-6, (ep, ev)
seen = torch.zeros(n_parts ** 3, dtype=torch.int64, device="cuda"); seen[i] += 1
dist.all_reduce(seen)
assert bool((seen == 1).all())
# and the slab step starts from them
ranks = pm.slab_ic.make_ranks_from_ic(comm, device=local)
tot = torch.tensor([ranks[0].count], device="cuda"); dist.all_reduce(tot)
assert int(tot.item()) == n_parts ** 3
a, da = 0.01, 0.0099
for s in range(3):
    pm.step(pos, vel, a, da, mass=8.0)
    pm.slab.slab_step(ranks, comm, a, da, mass=8.0, cfg=cfg)
    a += da
got_p, got_v = pm.slab.collect(ranks, comm, n_parts ** 3)
dd = torch.remainder(got_p.double() - pos.double() + n_cells / 2, n_cells) - n_cells / 2
ep2 = float(dd.norm() / pos.double().norm()); ev2 = float((got_v.double() - vel.double()).norm() / vel.double().norm())
assert ep2 <= 1e-5 and ev2 <= 1e-5, (ep2, ev2)
for r in ranks: r.close()
dist.barrier(); dist.destroy_process_group()
sys.stdout.write("rank" + str(comm.rank) + "-ok %%.2e %%.2e %%.2e %%.2e\n" %% (ep, ev, ep2, ev2)); sys.stdout.flush()
'''


def test_two_rank_nccl_slab_initial_conditions(tmp_path):
    """slab_ic over NCCL (all_to_all_single transposes of the complex128 fields, all-to-all-v routing)
    against the single-GPU generator, then three slab steps from those particles."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker_ic.py"
    script.write_text(IC_WORKER % REPO)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("-ok") == 2, out.stdout
