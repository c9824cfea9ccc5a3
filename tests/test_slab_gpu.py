"""Slab-decomposed step checked on ONE GPU by looping the ranks: P plans on the
same device, exchanges by tensor copies (slab.LocalComm).  The kernels, buffers and call sequence
are exactly those of the multi-process NCCL run; only the transport differs."""
import types

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cosmological_particle_mesh_simulation_b200 as pm
    yield pm
    pm.set_config(None)
    pm.release_plans()


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def rel_periodic(a, b, n):
    d = torch.remainder(a.double() - b.double() + n / 2, n) - n / 2
    return float(d.norm() / b.double().norm())


def test_particle_to_slab_assignment_bit_exact(pm):
    n, P = 64, 4
    pos, _ = O.lattice_ic(16, n, seed=3)
    pos[2, :5] = [64.0, 0.0, 15.999999, 16.0, 63.99999]       # Q4: z == Nc belongs to plane 0 -> rank 0
    want = (np.floor(pos[2]).astype(np.int64) % n) // (n // P)
    got_np = pm.slab.slab_of_particles(pos[2], n, P)
    got_t = pm.slab.slab_of_particles(torch.from_numpy(pos[2]).cuda(), n, P).cpu().numpy()
    assert np.array_equal(got_np, want) and np.array_equal(got_t, want)
    assert want[0] == 0 and want[2] == 0 and want[3] == 1 and want[4] == 3
    keys = O.cell_keys(pos, O.Config(N_CELLS=n))
    assert np.array_equal(want, (keys // (n * n)) // (n // P))  # same rule as the cell key


@pytest.mark.parametrize("transport", ["nccl", "peer", "fused"])
@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("n_parts,n_cells", [(32, 64), (64, 128)])
def test_slab_steps_equal_single_gpu_steps(pm, P, n_parts, n_cells, transport):
    cfg = types.SimpleNamespace(**O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100).__dict__)
    pm.set_config(cfg)
    if n_cells // P < 16:
        pytest.skip("slab thinner than one FFT column tile")
    pos_h, vel_h = O.lattice_ic(n_parts, n_cells, seed=11, vel_rms=0.3)
    # a few particles parked right at slab boundaries and at z == Nc so that migration, the ghost
    # planes and Q4 are exercised from step 1
    nzl = n_cells // P
    pos_h[2, :4] = [n_cells, nzl - 1e-3, nzl + 1e-3, n_cells - 1e-3]
    vel_h[2, :4] = [0.5, 2.0, -2.0, 3.0]
    pos, vel = torch.from_numpy(pos_h).cuda(), torch.from_numpy(vel_h).cuda()
    npart = pos.shape[1]
    mass = (n_cells / n_parts) ** 3

    comm = pm.slab.LocalComm(P)
    ranks = pm.slab.make_ranks(n_cells, pos, vel, comm)
    assert sum(r.count for r in ranks) == npart
    if transport != "nccl":      # transposes by stores into / loads from the other ranks' buffers
        assert pm.slab.setup_peers(ranks, comm)
    ref_p, ref_v = pos.clone(), vel.clone()
    moved = 0
    a, da = 0.3, 0.0099
    for step in range(6):
        rho_ref = torch.empty((n_cells,) * 3, device="cuda")
        pm.step(ref_p, ref_v, a, da, mass=mass, rho_out=rho_ref)
        before = [r.count for r in ranks]
        chunks = (1, 2, 4)[step % 3]            # unchunked, two and four kx chunks
        if chunks == 4 and (n_cells // 2 // 4) % 16:
            chunks = 2
        if chunks == 2 and (n_cells // 2 // 2) % 16:
            chunks = 1
        pm.slab.slab_step(ranks, comm, a, da, mass=mass, cfg=cfg, chunks=chunks, transport=transport)
        moved += sum(abs(x - r.count) for x, r in zip(before, ranks))
        assert sum(r.count for r in ranks) == npart            # nobody lost in migration
        rho = torch.cat([r.buf["RHO"] for r in ranks])          # density the step just used
        phi = torch.cat([r.buf["PHI"] for r in ranks])
        assert rel(rho, rho_ref) <= 1e-6, f"density step {step}"
        got_p, got_v = pm.slab.collect(ranks, comm, npart)
        assert rel_periodic(got_p, ref_p, n_cells) <= 1e-6, f"positions step {step}"
        assert rel(got_v, ref_v) <= 1e-5, f"velocities step {step}"
        # every particle sits on the rank that owns its z cell
        for r in ranks:
            p, _, _ = r.export()
            own = pm.slab.slab_of_particles(p[2], n_cells, P)
            assert bool((own == r.rank).all())
        a += da
    assert torch.isfinite(phi).all()
    if transport != "nccl":
        assert all(r.peer_timeouts() == 0 for r in ranks)
    for r in ranks:
        r.close()


def test_peer_and_nccl_transports_are_bit_identical(pm):
    """The three transports move the same numbers: the whole state must agree bit for bit."""
    n_parts, n_cells, P = 64, 128, 4
    cfg = types.SimpleNamespace(**O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100).__dict__)
    pm.set_config(cfg)
    pos_h, vel_h = O.lattice_ic(n_parts, n_cells, seed=5, vel_rms=0.5)
    outs = []
    for transport in ("nccl", "peer", "fused"):
        pos, vel = torch.from_numpy(pos_h).cuda(), torch.from_numpy(vel_h).cuda()
        comm = pm.slab.LocalComm(P)
        ranks = pm.slab.make_ranks(n_cells, pos, vel, comm)
        if transport != "nccl":
            assert pm.slab.setup_peers(ranks, comm)
        for s in range(4):
            pm.slab.slab_step(ranks, comm, 0.4 + 0.0099 * s, 0.0099, mass=8.0, cfg=cfg, chunks=(1, 2, 4, 2)[s],
                              transport=transport)
        phi = torch.cat([r.buf["PHI"] for r in ranks]).clone()
        outs.append(pm.slab.collect(ranks, comm, pos.shape[1]) + (phi,))
        for r in ranks:
            r.close()
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert torch.equal(x, y)


IPC_WORKER = r'''
import os, sys, types, torch, torch.distributed as dist
sys.path.insert(0, %r)
import cosmological_particle_mesh_simulation_b200 as pm
from oracle import oracle as O
torch.cuda.set_device(0)                       # both processes share the one GPU
dist.init_process_group("gloo")                # control plane only; data exchanges staged through the host
n_parts, n_cells = 32, 64
cfg = types.SimpleNamespace(**O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100).__dict__)
pm.set_config(cfg)
pos_h, vel_h = O.lattice_ic(n_parts, n_cells, seed=11, vel_rms=0.3)
pos, vel = torch.from_numpy(pos_h).cuda(), torch.from_numpy(vel_h).cuda()
comm = pm.slab.DistComm()
ranks = pm.slab.make_ranks(n_cells, pos, vel, comm, device=0)
assert pm.slab.setup_peers(ranks, comm), "CUDA IPC mapping / flag handshake failed"
assert pm.slab.setup_ghost_peers(ranks, comm), "mapping of the ghost / migration buffers failed"
ref_p, ref_v = pos.clone(), vel.clone()
a, da = 0.3, 0.0099
for s in range(3):
    pm.step(ref_p, ref_v, a, da, mass=8.0)
    pm.slab.slab_step(ranks, comm, a, da, mass=8.0, cfg=cfg, chunks=(1, 2, 1)[s], transport=("fused", "peer", "fused")[s],
                      ghosts=("peer", "peer", "nccl")[s], migrate=("peer", "nccl", "peer")[s])
    a += da
torch.cuda.synchronize()
assert ranks[0].peer_timeouts() == 0
p, v, ids = ranks[0].export()
d = torch.remainder(p.double() - ref_p[:, ids.long()].double() + n_cells / 2, n_cells) - n_cells / 2
ep = float(d.norm() / ref_p.double().norm()); ev = float((v.double() - ref_v[:, ids.long()].double()).norm() / ref_v.double().norm())
assert ep <= 1e-6 and ev <= 1e-5, (ep, ev)
tot = torch.tensor([ranks[0].count]); dist.all_reduce(tot)
assert int(tot.item()) == pos.shape[1]
pm.slab.release_peers(ranks, comm)
for r in ranks: r.close()
dist.barrier(); dist.destroy_process_group()
sys.stdout.write("rank" + str(comm.rank) + "-ok %%.2e %%.2e\n" %% (ep, ev)); sys.stdout.flush()
'''


def test_peer_transport_through_cuda_ipc_two_processes_one_gpu(tmp_path):
    """Two ranks = two processes on the SAME GPU: the z-pass arrays and flag words of the other rank
    are reached through cudaIpcOpenMemHandle mappings exactly as between two GPUs of a box."""
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(IPC_WORKER % repo)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("-ok") == 2, out.stdout


@pytest.mark.parametrize("P", [1, 2, 4])
def test_ghost_planes_and_migration_through_peer_memory_are_bit_identical(pm, P):
    """Ghost planes pushed into the neighbours' memory, particle migration through peer memory (count
    matrix + records, csrc/pm_migrate.cu) and the two-stream fused schedule move the same numbers as
    the NCCL send/recv and all-to-all-v routes: particles (by id) and potential bit for bit."""
    n_parts, n_cells = 64, 128
    cfg = types.SimpleNamespace(**O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100).__dict__)
    pm.set_config(cfg)
    pos_h, vel_h = O.lattice_ic(n_parts, n_cells, seed=5, vel_rms=0.5)
    outs = []
    for transport, ghosts, migrate in (("fused", "nccl", "nccl"), ("fused", "peer", "nccl"), ("fused", "peer", "peer"),
                                       ("fused2", "peer", "peer"), ("nccl", "peer", "peer"), ("fused", None, None)):
        pos, vel = torch.from_numpy(pos_h).cuda(), torch.from_numpy(vel_h).cuda()
        comm = pm.slab.LocalComm(P)
        ranks = pm.slab.make_ranks(n_cells, pos, vel, comm)
        assert pm.slab.setup_peers(ranks, comm) and pm.slab.setup_ghost_peers(ranks, comm)
        moved = 0
        for s in range(4):
            before = [r.count for r in ranks]
            pm.slab.slab_step(ranks, comm, 0.4 + 0.0099 * s, 0.0099, mass=8.0, cfg=cfg, chunks=(1, 2, 2, 1)[s],
                              transport=transport, ghosts=ghosts, migrate=migrate)
            moved += sum(abs(x - r.count) for x, r in zip(before, ranks))
        assert all(r.peer_timeouts() == 0 for r in ranks)
        assert sum(r.count for r in ranks) == pos.shape[1]
        if P > 1:
            assert moved > 0                      # the migration route really carried particles
        phi = torch.cat([r.buf["PHI"] for r in ranks]).clone()
        outs.append(pm.slab.collect(ranks, comm, pos.shape[1]) + (phi,))
        for r in ranks:
            r.close()
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert torch.equal(x, y)


def test_exchange_errors_are_raised_by_every_rank(pm):
    """A leave list that overflows on ONE rank is seen in the count matrix by all of them: every rank
    raises SlabExchangeError in the same step (nobody is left waiting in a collective)."""
    n_parts, n_cells, P = 32, 64, 2
    cfg = types.SimpleNamespace(**O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100).__dict__)
    pm.set_config(cfg)
    pos_h, vel_h = O.lattice_ic(n_parts, n_cells, seed=5, vel_rms=0.1)
    pos, vel = torch.from_numpy(pos_h).cuda(), torch.from_numpy(vel_h).cuda()
    comm = pm.slab.LocalComm(P)
    ranks = pm.slab.make_ranks(n_cells, pos, vel, comm)
    assert pm.slab.setup_peers(ranks, comm) and pm.slab.setup_ghost_peers(ranks, comm)
    pm.slab.slab_step(ranks, comm, 0.4, 0.0099, mass=8.0, cfg=cfg)
    # fake an overflow on rank 1: its leave count towards rank 0 beyond the list capacity
    ranks[1].buf["LEAVE_COUNTS"][0] = 2 ** 30
    for r in ranks:
        r.migrate_counts_push()
    mats = [r.migrate_counts_read() for r in ranks]
    assert all((m == mats[0]).all() for m in mats) and mats[0][1, P + 1] == 1 and mats[0][0, P + 1] == 0
    # the NCCL migration path: the same overflow, agreed on by a flag all-reduce before the all-to-all-v
    pm.slab.slab_step(ranks, comm, 0.41, 0.0099, mass=8.0, cfg=cfg, migrate="nccl")
    real_gather = pm.slab.SlabRank.gather

    def gather_then_overflow(self, *args):
        real_gather(self, *args)
        if self.rank == 1:
            self.buf["LEAVE_COUNTS"][0] = 2 ** 30
    pm.slab.SlabRank.gather = gather_then_overflow
    try:
        with pytest.raises(pm.slab.SlabExchangeError):
            pm.slab.slab_step(ranks, comm, 0.42, 0.0099, mass=8.0, cfg=cfg, migrate="nccl")
    finally:
        pm.slab.SlabRank.gather = real_gather
    for r in ranks:
        r.close()


def test_slab_run_is_deterministic(pm):
    n_parts, n_cells, P = 32, 64, 2
    cfg = types.SimpleNamespace(**O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100).__dict__)
    pm.set_config(cfg)
    pos_h, vel_h = O.lattice_ic(n_parts, n_cells, seed=2, vel_rms=1.0)
    outs = []
    for _ in range(2):
        pos, vel = torch.from_numpy(pos_h).cuda(), torch.from_numpy(vel_h).cuda()
        comm = pm.slab.LocalComm(P)
        ranks = pm.slab.make_ranks(n_cells, pos, vel, comm)
        for s in range(4):
            pm.slab.slab_step(ranks, comm, 0.5 + 0.0099 * s, 0.0099, mass=8.0, cfg=cfg)
        outs.append(pm.slab.collect(ranks, comm, pos.shape[1]))
        for r in ranks:
            r.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("n_cells,nranks", [(512, 2), (1024, 8), (2048, 8)])
def test_slab_row_passes_two_stage_equal_radix8(pm, n_cells, nranks, monkeypatch):
    """The slab's row passes run the two-stage register-resident kernels (pm_fft2.cuh; 2048 points: rows only)
    where they exist and the radix-8 kernels otherwise (PM_FFT_V2=0 forces them).  Same layout and Nyquist
    packing, so forward + inverse over one rank's planes must agree to float32 rounding, and both must return
    2 * Nc/2 ... i.e. Nc times the (mean-free) input.  One rank of an `nranks` geometry is enough: the row
    passes are local.  (2048: 256 planes of 2048^2 = 4.3 GB per array.)"""
    slab = pm.slab
    rng = torch.Generator(device="cuda").manual_seed(5)
    out = {}
    x = None
    for v2 in ("1", "0"):
        monkeypatch.setenv("PM_FFT_V2", v2)
        r = slab.SlabRank(n_cells, 4096, torch.cuda.current_device(), 0, nranks)
        try:
            if x is None:
                x = torch.rand(r.buf["RHO"].shape, generator=rng, device="cuda", dtype=torch.float32)
            r.buf["RHO"].copy_(x)
            pm._runtime.check(pm._runtime.lib().pm_slab_set_rho_mean(r.handle, 0.5), "pm_slab_set_rho_mean")
            r.fft_rows_forward()
            r.fft_rows_inverse()
            torch.cuda.synchronize()
            out[v2] = r.buf["PHI"].clone()
        finally:
            r.close()
    a, b = out["1"], out["0"]
    want = (x - 0.5) * float(n_cells)
    scale = float(want.double().norm())
    assert float((a.double() - b.double()).norm()) / scale <= 2e-6
    assert float((a.double() - want.double()).norm()) / scale <= 2e-6
    del out, a, b, want, x
    torch.cuda.empty_cache()
