"""Initial conditions generated slab by slab (SURVEY 8f row f1 for BASELINE configs[2]-[3]).

The reference builds its initial conditions in one address space: gaussian_random_field()
(src/gaussian_random_field.py:9-29) and zeldovich(density) (src/zeldovich.py:10-100), three
complex N_PARTS^3 transforms deep.  At 1024^3 particles on eight GPUs no rank should have to hold
the whole lattice, so here every rank holds 1/P of every field and the 3-D transforms are split
around an all-to-all transpose:

    k-space fields      planes  [r n/P, (r+1) n/P) of the slowest array axis   ([n/P][n][n])
    real-space fields   columns [r n/P, (r+1) n/P) of the fastest array axis   ([n][n][n/P])

    inverse:  2-D transform of the planes  -> transpose -> 1-D transform along axis 0
    forward:  2-D transform over axes 0, 1 -> transpose -> 1-D transform along axis 2

The fastest lattice axis is the one that becomes positions[2] (zeldovich.py:83, meshgrid 'ij'),
i.e. the axis the mesh slabs are cut along (slab.py), so a rank generates the particles whose
lattice point lies in its own mesh slab; the few that the displacement and the +-2 cell jitter carry
across a slab boundary are routed to their owner once (`route_to_owners`, one all-to-all-v).

All arithmetic is in libpmstep.so (pm_ic_slab_* / pm_ic_noise_range of include/pmstep.h: Philox noise
and jitter as functions of the GLOBAL element / particle index, cuFFT Z2Z for the partial
transforms); this module sequences the calls and does the transposes through a slab.Comm
(torch.distributed all_to_all_single over NCCL, or LocalComm with all ranks on one GPU).  Particle
ids are the single-GPU generator's particle indices, and the union of the slabs equals
gaussian_random_field() + zeldovich() of this package up to the rounding of the re-ordered
transforms (tests/test_slab_ic.py: <= 1e-6, almost every value bit-identical).
"""
from __future__ import annotations

import torch

try:
    from . import _runtime as rt
    from .slab import slab_of_particles, make_rank_from_local
except ImportError:  # flat layout
    import _runtime as rt
    from slab import slab_of_particles, make_rank_from_local

AXES_12, AXIS_0, AXES_01, AXIS_2 = 0, 1, 2, 3   # PM_IC_AXES_* of include/pmstep.h


class DeviceOps:
    """The arithmetic of the generator: thin calls into the C ABI on one CUDA device."""

    def __init__(self, cfg, device, seed):
        self.cfg, self.dev, self.seed = cfg, int(device), int(seed)
        self.prm = rt.ic_params(cfg)
        self.n = int(cfg.N_PARTS)
        self._lib = rt.lib()
        nbytes = int(self._lib.pm_ic_slab_workspace_bytes())
        self._work = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{self.dev}")
        self._work_bytes = nbytes

    def _st(self):
        return rt.stream_ptr(self.dev)

    def _z(self, shape):
        return torch.empty(shape, dtype=torch.complex128, device=f"cuda:{self.dev}")

    def rho_k(self, i0_lo, n0l):
        """sqrt(p D^2) (f1 + i f2) on planes i0_lo .. i0_lo + n0l - 1 (gaussian_random_field.py:14-24)."""
        n = self.n
        cnt = n0l * n * n
        f1 = torch.empty(cnt, dtype=torch.float32, device=f"cuda:{self.dev}")
        f2 = torch.empty_like(f1)
        z = self._z((n0l, n, n))
        with torch.cuda.device(self.dev):
            rt.check(self._lib.pm_ic_noise_range(f1.data_ptr(), f2.data_ptr(), i0_lo * n * n, cnt, self.seed, self._st()),
                     "pm_ic_noise_range")
            rt.check(self._lib.pm_ic_slab_rho_k(self.prm, f1.data_ptr(), f2.data_ptr(), i0_lo, n0l, z.data_ptr(),
                                                self._work.data_ptr(), self._work_bytes, self._st()), "pm_ic_slab_rho_k")
        return z

    def fft(self, z, axes, inverse):
        d0, d1, d2 = z.shape
        with torch.cuda.device(self.dev):
            rt.check(self._lib.pm_ic_slab_fft(z.data_ptr(), d0, d1, d2, axes, 1 if inverse else 0, self._st()),
                     "pm_ic_slab_fft")
        return z

    def real_f32(self, z, scale):
        out = torch.empty(z.shape, dtype=torch.float32, device=z.device)
        with torch.cuda.device(self.dev):
            rt.check(self._lib.pm_ic_slab_real_f32(z.data_ptr(), z.numel(), float(scale), out.data_ptr(), self._st()),
                     "pm_ic_slab_real_f32")
        return out

    def from_f32(self, x):
        z = self._z(tuple(x.shape))
        with torch.cuda.device(self.dev):
            rt.check(self._lib.pm_ic_slab_from_f32(x.data_ptr(), x.numel(), z.data_ptr(), self._st()), "pm_ic_slab_from_f32")
        return z

    def displacement_k(self, direction, rho_k, i0_lo):
        out = torch.empty_like(rho_k)
        with torch.cuda.device(self.dev):
            rt.check(self._lib.pm_ic_slab_displacement_k(self.prm, direction, rho_k.data_ptr(), i0_lo, rho_k.shape[0],
                                                         out.data_ptr(), self._st()), "pm_ic_slab_displacement_k")
        return out

    def particles(self, direction, z, i2_lo, pos_row, vel_row, ids):
        with torch.cuda.device(self.dev):
            rt.check(self._lib.pm_ic_slab_particles(self.prm, direction, z.data_ptr(), i2_lo, z.shape[2], self.seed, None,
                                                    pos_row.data_ptr(), vel_row.data_ptr(),
                                                    ids.data_ptr() if ids is not None else None, self._st()),
                     "pm_ic_slab_particles")

    def empty_particles(self, cnt):
        dev = f"cuda:{self.dev}"
        return (torch.empty((3, cnt), dtype=torch.float32, device=dev), torch.empty((3, cnt), dtype=torch.float32, device=dev),
                torch.empty(cnt, dtype=torch.int32, device=dev))


# ---- the two transposes (pure data movement; any complex tensors, any Comm) -------------------------

def planes_to_columns(planes, comm):
    """Per local rank [n/P][n][n] (planes of axis 0) -> [n][n][n/P] (columns of axis 2)."""
    P = comm.nranks
    send, recv = [], []
    for a in planes:
        nl, n, n2 = a.shape
        assert n2 == n and nl * P == n, "planes_to_columns: expected [n/P][n][n]"
        blk = a.view(nl, n, P, nl).permute(2, 0, 1, 3).contiguous()          # [dst][i0 local][i1][i2 local]
        send.append(torch.view_as_real(blk))
        recv.append(torch.empty_like(send[-1]))
    comm.all_to_all(send, recv)
    # recv[src][i0 local][i1][i2 local]: (src, i0 local) is the global i0
    return [torch.view_as_complex(r).view(r.shape[0] * r.shape[1], r.shape[2], r.shape[3]) for r in recv]


def columns_to_planes(cols, comm):
    """Per local rank [n][n][n/P] (columns of axis 2) -> [n/P][n][n] (planes of axis 0)."""
    P = comm.nranks
    send, recv = [], []
    for c in cols:
        n, n1, nl = c.shape
        assert n1 == n and nl * P == n, "columns_to_planes: expected [n][n][n/P]"
        send.append(torch.view_as_real(c.contiguous().view(P, nl, n, nl)))   # [dst][i0 local][i1][i2 local]
        recv.append(torch.empty_like(send[-1]))
    comm.all_to_all(send, recv)
    out = []
    for r in recv:                                                           # [src][i0 local][i1][i2 local of src]
        z = torch.view_as_complex(r)
        Pn, nl, n, _ = z.shape
        out.append(z.permute(1, 2, 0, 3).contiguous().view(nl, n, Pn * nl))
    return out


def _split(n, P):
    if n % P:
        raise ValueError(f"slab initial conditions: N_PARTS = {n} is not a multiple of the {P} ranks")
    return n // P


def slab_gaussian_random_field(comm, ops):
    """gaussian_random_field() (gaussian_random_field.py:9-29) with the field held as columns:
    per local rank the float32 density [n][n][n/P] and nothing else."""
    P = comm.nranks
    nl = _split(ops[0].n, P)
    n = ops[0].n
    planes = [o.fft(o.rho_k(r * nl, nl), AXES_12, True) for o, r in zip(ops, comm.local_ranks)]
    cols = planes_to_columns(planes, comm)
    del planes
    return [o.real_f32(o.fft(c.contiguous(), AXIS_0, True), 1.0 / float(n) ** 3) for o, c in zip(ops, cols)]


def slab_zeldovich(density_cols, comm, ops):
    """zeldovich(density) (zeldovich.py:10-22) on column-decomposed fields: per local rank
    (positions [3, n n n/P], velocities, ids) of the lattice points in its columns, in the local
    order (i0, i1, i2 local); ids are the single-GPU particle indices (i0 n + i1) n + i2."""
    P = comm.nranks
    n = ops[0].n
    nl = _split(n, P)
    # np.fft.fftn(density), zeldovich.py:17
    cols = [o.fft(o.from_f32(d), AXES_01, False) for o, d in zip(ops, density_cols)]
    rho_k = [o.fft(z, AXIS_2, False) for o, z in zip(ops, columns_to_planes(cols, comm))]
    del cols
    out = [o.empty_particles(n * n * nl) for o in ops]
    for direction in (0, 1, 2):
        planes = [o.fft(o.displacement_k(direction, z, r * nl), AXES_12, True)
                  for o, z, r in zip(ops, rho_k, comm.local_ranks)]
        cols = planes_to_columns(planes, comm)
        del planes
        for o, c, r, (pos, vel, ids) in zip(ops, cols, comm.local_ranks, out):
            z = o.fft(c.contiguous(), AXIS_0, True)
            o.particles(direction, z, r * nl, pos[direction], vel[direction], ids if direction == 0 else None)
        del cols
    return out


def route_to_owners(parts, comm, n_cells):
    """Send every particle to the rank whose mesh slab holds its z cell (slab.slab_of_particles).
    parts: per local rank (pos [3, k], vel [3, k], ids int32 [k]).  Returns the same triples with
    the arrivals appended source rank by source rank (deterministic order)."""
    P = comm.nranks
    send, send_counts = [], []
    for pos, vel, ids in parts:
        owner = slab_of_particles(pos[2], n_cells, P).to(torch.int64)
        order = torch.argsort(owner, stable=True)
        rec = torch.empty((pos.shape[1], 7), dtype=torch.float32, device=pos.device)
        rec[:, 0:3] = pos.t()[order]
        rec[:, 3:6] = vel.t()[order]
        rec[:, 6] = ids[order].view(torch.float32)
        send.append(rec)
        send_counts.append(torch.bincount(owner, minlength=P).tolist())
    recv_counts = comm.exchange_counts(send_counts)
    recv = [torch.empty((sum(int(c) for c in rc), 7), dtype=torch.float32, device=s.device)
            for rc, s in zip(recv_counts, send)]
    comm.all_to_all_v(send, send_counts, recv, recv_counts)
    out = []
    for r in recv:
        t = r.t().contiguous()
        out.append((t[0:3].contiguous(), t[3:6].contiguous(), t[6].contiguous().view(torch.int32)))
    return out


def slab_initial_conditions(comm, cfg=None, device=None, seed=None, route=True, ops=None, return_density=False):
    """The package's gaussian_random_field() + zeldovich() for a slab-decomposed run.
    Returns per local rank (positions [3, k], velocities [3, k], ids int32 [k]); with `route`
    (default) every particle is on the rank that owns its z cell.  return_density: also the
    float32 Gaussian field, per local rank its columns [n][n][n/P] (the mesh the t = 0 snapshot holds)."""
    cfg = cfg if cfg is not None else rt.config()
    if ops is None:
        dev = rt.current_device() if device is None else int(device)
        seed = int(cfg.RANDOM_SEED if seed is None else seed)
        ops = [DeviceOps(cfg, dev, seed) for _ in comm.local_ranks]
    density = slab_gaussian_random_field(comm, ops)
    parts = slab_zeldovich(density, comm, ops)
    if route:
        parts = route_to_owners(parts, comm, int(cfg.N_CELLS))
    return (parts, density) if return_density else parts


def make_ranks_from_ic(comm, cfg=None, device=None, seed=None, slack=1.25, return_density=False):
    """slab.SlabRank(s) of this process loaded with slab-generated initial conditions."""
    cfg = cfg if cfg is not None else rt.config()
    dev = rt.current_device() if device is None else int(device)
    total = int(cfg.N_PARTS) ** 3
    cap = int(total / comm.nranks * slack) + 4096
    parts, density = slab_initial_conditions(comm, cfg=cfg, device=dev, seed=seed, return_density=True)
    ranks = [make_rank_from_local(int(cfg.N_CELLS), p, v, i, r, comm.nranks, device=dev, capacity=cap,
                                  total_particles=total)
             for (p, v, i), r in zip(parts, comm.local_ranks)]
    return (ranks, density) if return_density else ranks
