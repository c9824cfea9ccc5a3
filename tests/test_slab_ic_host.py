"""Host logic of the slab-by-slab initial conditions (slab_ic.py) on CPU: the two transposes, the
axis split of the 3-D transforms, the particle order / ids of a column slab and the routing to the
owners -- with the device arithmetic replaced by a TEST DOUBLE built on the IC oracle
(oracle/oracle_ic.py, NumPy), so that the union of the slabs must reproduce the oracle's
single-address-space zeldovich() to rounding.  The GPU test (tests/test_slab_ic.py) runs the same
module with the real C-ABI calls."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import oracle_ic as IC  # noqa: E402
from cosmological_particle_mesh_simulation_b200 import slab_ic  # noqa: E402
from cosmological_particle_mesh_simulation_b200.slab import LocalComm, slab_of_particles  # noqa: E402


class OracleOps:
    """slab_ic.DeviceOps with every device call replaced by the oracle's NumPy statement of the same
    piece (test infrastructure only)."""

    def __init__(self, cfg, f1, f2, jitter):
        self.cfg, self.n = cfg, cfg.N_PARTS
        self.f1, self.f2, self.jitter = f1, f2, jitter

    def rho_k(self, i0_lo, n0l):
        D = IC.Dt(self.cfg.A_INIT, [self.cfg.OMEGA_M0, self.cfg.OMEGA_LAMBDA0, self.cfg.OMEGA_K0])
        amp = np.sqrt(IC.power_spectrum(self.cfg) * D ** 2)
        z = amp * self.f1 + 1j * (amp * self.f2)
        return torch.from_numpy(np.ascontiguousarray(z[i0_lo:i0_lo + n0l]))

    def fft(self, z, axes, inverse):
        dims = {slab_ic.AXES_12: (1, 2), slab_ic.AXIS_0: (0,), slab_ic.AXES_01: (0, 1), slab_ic.AXIS_2: (2,)}[axes]
        out = torch.fft.ifftn(z, dim=dims, norm="forward") if inverse else torch.fft.fftn(z, dim=dims)  # unnormalised
        z.copy_(out)
        return z

    def real_f32(self, z, scale):
        return (z.real * scale).to(torch.float32)

    def from_f32(self, x):
        return x.to(torch.complex128)

    def displacement_k(self, direction, rho_k, i0_lo):
        n = self.n
        full = np.zeros((n, n, n), dtype=np.complex128)
        n0l = rho_k.shape[0]
        full[i0_lo:i0_lo + n0l] = rho_k.numpy()
        d = IC.displacement_field_k(IC.potential_k(full, self.cfg), direction, self.cfg)
        return torch.from_numpy(np.ascontiguousarray(d[i0_lo:i0_lo + n0l]))

    def empty_particles(self, cnt):
        return torch.empty((3, cnt), dtype=torch.float32), torch.empty((3, cnt), dtype=torch.float32), \
            torch.empty(cnt, dtype=torch.int32)

    def particles(self, direction, z, i2_lo, pos_row, vel_row, ids):
        n, n2l = self.n, z.shape[2]
        i0, i1, i2 = np.meshgrid(np.arange(n), np.arange(n), np.arange(i2_lo, i2_lo + n2l), indexing="ij")
        gid = ((i0 * n + i1) * n + i2).reshape(-1)
        disp_local = (z.numpy().real / float(n) ** 3).reshape(-1) * (self.cfg.N_CELLS / self.cfg.BOX_SIZE)
        # the oracle's formulas work on the whole lattice: scatter the slab into it, take the slab back
        disp = np.zeros(n ** 3)
        disp[gid] = disp_local
        pos = IC.zeldovich_positions(disp, direction, self.jitter[direction], self.cfg)[gid]
        vel = IC.zeldovich_velocities(disp, self.cfg)[gid]
        pos_row.copy_(torch.from_numpy(pos.astype(np.float32)))
        vel_row.copy_(torch.from_numpy(vel.astype(np.float32)))
        if ids is not None:
            ids.copy_(torch.from_numpy(gid.astype(np.int32)))


@pytest.mark.parametrize("P", [1, 2, 4])
def test_transposes_are_inverse_and_place_every_element(P):
    n = 8
    nl = n // P
    full = torch.arange(n ** 3, dtype=torch.float64).reshape(n, n, n).to(torch.complex128) * (1 + 2j)
    comm = LocalComm(P)
    planes = [full[r * nl:(r + 1) * nl].clone() for r in range(P)]
    cols = slab_ic.planes_to_columns(planes, comm)
    for r in range(P):
        assert cols[r].shape == (n, n, nl)
        assert torch.equal(cols[r], full[:, :, r * nl:(r + 1) * nl])
    back = slab_ic.columns_to_planes(cols, comm)
    for r in range(P):
        assert torch.equal(back[r], planes[r])


@pytest.mark.parametrize("P", [1, 2, 4])
def test_split_transforms_equal_fftn(P):
    n = 8
    rng = np.random.default_rng(5)
    full = torch.from_numpy(rng.standard_normal((n, n, n)) + 1j * rng.standard_normal((n, n, n)))
    ops = OracleOps(IC.ICConfig(N_PARTS=n, N_CELLS=2 * n), None, None, None)
    comm, nl = LocalComm(P), n // P
    planes = [ops.fft(full[r * nl:(r + 1) * nl].clone(), slab_ic.AXES_12, True) for r in range(P)]
    cols = [ops.fft(c.contiguous(), slab_ic.AXIS_0, True) for c in slab_ic.planes_to_columns(planes, comm)]
    want = torch.fft.ifftn(full, norm="forward")
    for r in range(P):
        assert torch.allclose(cols[r], want[:, :, r * nl:(r + 1) * nl], rtol=1e-12, atol=1e-12)
    fwd = [ops.fft(c.clone(), slab_ic.AXES_01, False) for c in cols]
    planes2 = [ops.fft(p, slab_ic.AXIS_2, False) for p in slab_ic.columns_to_planes(fwd, comm)]
    for r in range(P):
        assert torch.allclose(planes2[r], full[r * nl:(r + 1) * nl] * n ** 3, rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize("P", [1, 2, 4])
def test_slab_pipeline_reproduces_the_oracle_generator(P):
    cfg = IC.ICConfig(N_PARTS=8, N_CELLS=16)
    n, n3 = cfg.N_PARTS, cfg.N_PARTS ** 3
    rng = np.random.default_rng(11)
    f1 = rng.standard_normal((n, n, n)).astype(np.float32)
    f2 = rng.standard_normal((n, n, n)).astype(np.float32)
    jitter = rng.uniform(-2, 2, size=(3, n3))
    rho = IC.gaussian_random_field(f1, f2, cfg)
    pos_o, vel_o = IC.zeldovich(rho, jitter, cfg)

    comm = LocalComm(P)
    ops = [OracleOps(cfg, f1, f2, jitter) for _ in range(P)]
    dens = slab_ic.slab_gaussian_random_field(comm, ops)
    nl = n // P
    for r in range(P):
        np.testing.assert_allclose(dens[r].numpy(), rho[:, :, r * nl:(r + 1) * nl], rtol=0, atol=1e-6 * np.abs(rho).max())
    parts = slab_ic.slab_initial_conditions(comm, cfg=cfg, ops=ops)
    seen = np.zeros(n3, dtype=np.int64)
    for r, (pos, vel, ids) in enumerate(parts):
        i = ids.numpy().astype(np.int64)
        seen[i] += 1
        # every particle sits on the rank that owns its z cell
        assert (slab_of_particles(pos[2].numpy(), cfg.N_CELLS, P) == r).all()
        d = (pos.numpy().astype(np.float64) - pos_o[:, i] + cfg.N_CELLS / 2) % cfg.N_CELLS - cfg.N_CELLS / 2
        assert np.abs(d).max() <= 4e-6 * cfg.N_CELLS          # a float32 ulp or two at most
        np.testing.assert_allclose(vel.numpy(), vel_o[:, i], rtol=0, atol=1e-6 * np.abs(vel_o).max())
    assert (seen == 1).all()                                   # every lattice point exactly once


def test_route_to_owners_keeps_records_intact():
    P, n_cells = 4, 16
    rng = np.random.default_rng(3)
    comm = LocalComm(P)
    parts, allp = [], []
    base = 0
    for r in range(P):
        k = 50 + 7 * r
        pos = torch.from_numpy(rng.uniform(0, n_cells, size=(3, k)).astype(np.float32))
        pos[2, 0] = float(n_cells)                             # z == N_CELLS wraps to cell 0
        vel = torch.from_numpy(rng.standard_normal((3, k)).astype(np.float32))
        ids = torch.arange(base, base + k, dtype=torch.int32)
        base += k
        parts.append((pos, vel, ids))
        allp.append((pos.clone(), vel.clone(), ids.clone()))
    out = slab_ic.route_to_owners(parts, comm, n_cells)
    pos_all = torch.cat([p for p, _, _ in allp], dim=1)
    vel_all = torch.cat([v for _, v, _ in allp], dim=1)
    got = 0
    for r, (pos, vel, ids) in enumerate(out):
        i = ids.long()
        assert torch.equal(pos, pos_all[:, i]) and torch.equal(vel, vel_all[:, i])
        assert (slab_of_particles(pos[2], n_cells, P) == r).all()
        got += ids.numel()
    assert got == base
    assert sorted(torch.cat([i for _, _, i in out]).tolist()) == list(range(base))


def test_ranks_must_divide_the_lattice():
    cfg = IC.ICConfig(N_PARTS=6, N_CELLS=12)
    ops = [OracleOps(cfg, None, None, None) for _ in range(4)]
    with pytest.raises(ValueError):
        slab_ic.slab_gaussian_random_field(LocalComm(4), ops)


GLOO_WORKER = r'''
import sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
sys.path.insert(0, %r)
from cosmological_particle_mesh_simulation_b200 import slab, slab_ic
from oracle import oracle_ic as IC
from test_slab_ic_host import OracleOps
dist.init_process_group("gloo")
r, P = dist.get_rank(), dist.get_world_size()
cfg = IC.ICConfig(N_PARTS=8, N_CELLS=16)
n = cfg.N_PARTS
rng = np.random.default_rng(11)                     # the same draws on every rank
f1 = rng.standard_normal((n, n, n)).astype(np.float32)
f2 = rng.standard_normal((n, n, n)).astype(np.float32)
jitter = rng.uniform(-2, 2, size=(3, n ** 3))
# the in-process rank loop is the reference for the distributed run
loc = slab_ic.slab_initial_conditions(slab.LocalComm(P), cfg=cfg, ops=[OracleOps(cfg, f1, f2, jitter) for _ in range(P)])
mine = slab_ic.slab_initial_conditions(slab.DistComm(), cfg=cfg, ops=[OracleOps(cfg, f1, f2, jitter)])
assert len(mine) == 1
for a, b in zip(mine[0], loc[r]):
    assert torch.equal(a, b)
dist.barrier(); dist.destroy_process_group()
sys.stdout.write("rank" + str(r) + "-ok\n"); sys.stdout.flush()
'''


def test_slab_ic_over_gloo_world2_equals_the_rank_loop(tmp_path):
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % (repo, os.path.join(repo, "tests")))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29537", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("-ok") == 2, out.stdout
