"""CPU oracle for the particle-mesh step -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``cosmological_particle_mesh_simulation_b200``) never does and has no CPU fallback.

It restates the reference's per-step loop (``/root/reference/src/pmesh.py:56-63``) with the same
call signatures as the reference modules; the particle loops live in ``pm_oracle.c`` (plain C,
``-ffp-contract=off``), the 3-D complex128 transforms are ``scipy.fft`` (the reference calls
pyFFTW -- FFTW3, version unpinned, not installed here; both compute the exact DFT to ~1e-15,
SURVEY.md section 8c).  Parity status: pinned against ``tests/golden/*.npz`` which
``oracle/make_golden.py`` produced by running the reference's own source files.

Unlike the reference (numba freezes ``N_CELLS`` at first call) the oracle takes the
configuration as an explicit :class:`Config`, so one process can check several mesh sizes.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np
import scipy.fft

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile pm_oracle.c -> libpm_oracle.so with the committed Makefile."""
    so = os.path.join(_HERE, "libpm_oracle.so")
    src = os.path.join(_HERE, "pm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libpm_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        i64, f64, vp, ci = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
        lib.pmo_max_threads.restype = ci
        lib.pmo_density.argtypes = [vp, i64, ci, f64, vp, ci]
        lib.pmo_potential_k.argtypes = [vp, vp, i64, f64, f64, ci]
        lib.pmo_integrate.argtypes = [vp, vp, i64, ci, f64, f64, f64, vp, vp, ci]
        lib.pmo_cell_keys.argtypes = [vp, i64, ci, vp]
        for fn in (lib.pmo_density, lib.pmo_potential_k, lib.pmo_integrate, lib.pmo_cell_keys):
            fn.restype = None
        _LIB = lib
    return _LIB


def max_threads() -> int:
    return int(_lib().pmo_max_threads())


@dataclass
class Config:
    """The configure_me.py names the hot path reads (configure_me.py:7-40)."""
    N_CELLS: int = 512
    N_PARTS: int = 256
    N_CPU: int = 1
    OMEGA_M0: float = 0.31
    OMEGA_K0: float = 0.00
    OMEGA_LAMBDA0: float = 0.69
    H0: float = 0.68
    A_INIT: float = 0.01
    A_END: float = 1.00
    STEPS: int = 1000


def _f32c(a, shape=None):
    assert a.dtype == np.float32 and a.flags.c_contiguous, "float32 C-contiguous expected"
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)
    return a.ctypes.data


def f(a, cosmology):
    """cosmology.py:20-27.  NB the loop calls it as f(a+da, [H0, OMEGA_LAMBDA0, OMEGA_K0])
    (integrate.py:12)."""
    omegaM, omegaL, omegaK = cosmology[0], cosmology[1], cosmology[2]
    return 1 / np.sqrt((omegaM + omegaK * a + omegaL * a ** 3) / a)


def density(positions, mass, cfg: Config):
    """density.py:7-48.  cfg.N_CPU == 1 gives the deterministic particle-order result."""
    npart = positions.shape[1]
    grid = np.empty((cfg.N_CELLS,) * 3, dtype=np.float32)
    _lib().pmo_density(_f32c(positions, (3, npart)), npart, cfg.N_CELLS, float(mass),
                       grid.ctypes.data, int(cfg.N_CPU))
    return grid


def fourier_grid(cfg: Config):
    """fourier_utils.py:5-16, with the uninitialised DC entry pinned to 0."""
    scale = 2 * np.pi
    k_x = np.array(scale * np.fft.fftfreq(cfg.N_CELLS), dtype='float32')
    k_y = np.array(scale * np.fft.fftfreq(cfg.N_CELLS), dtype='float32')
    k_z = np.array(scale * np.fft.fftfreq(cfg.N_CELLS), dtype='float32')
    ky, kz, kx = np.meshgrid(k_z, k_y, k_x)
    k_squared = np.sin(kz / 2) ** 2 + np.sin(ky / 2) ** 2 + np.sin(kx / 2) ** 2
    return np.divide(1, k_squared, out=np.zeros_like(k_squared), where=k_squared != 0)


def density_k(rho, cfg: Config):
    """potential.py:17-21: astype(cdouble) then an unnormalised forward c2c DFT."""
    return scipy.fft.fftn(rho.astype(np.cdouble), axes=(0, 1, 2), workers=int(cfg.N_CPU))


def potential_k(rho_k, fgrid, a, cfg: Config):
    """potential.py:12-15 (returns a new array, like the reference's array expression)."""
    out = np.ascontiguousarray(rho_k, dtype=np.cdouble).copy()
    fg = np.ascontiguousarray(fgrid, dtype=np.float32)
    _lib().pmo_potential_k(out.ctypes.data, fg.ctypes.data, out.size, float(cfg.OMEGA_M0),
                           float(a), int(cfg.N_CPU))
    return out


def potential_real(pot_k, cfg: Config):
    """potential.py:23-29: backward DFT normalised by 1/Nc^3 (pyFFTW default), real part, f32."""
    return (scipy.fft.ifftn(pot_k, axes=(0, 1, 2), workers=int(cfg.N_CPU)).real).astype('float32')


def potential(rho, fgrid, a, cfg: Config):
    """potential.py:7-10."""
    return potential_real(potential_k(density_k(rho, cfg), fgrid, a, cfg), cfg)


def integrate(positions, velocities, a_val, f_a1, da, potentials, cfg: Config, acc=None):
    """integrate.py:15-25; in place, returns the same arrays.  acc: optional float64[3,Np]
    receiving g_p (integrate.py:92) per direction."""
    npart = positions.shape[1]
    if acc is not None:
        assert acc.dtype == np.float64 and acc.shape == (3, npart) and acc.flags.c_contiguous
    _lib().pmo_integrate(_f32c(positions, (3, npart)), _f32c(velocities, (3, npart)), npart,
                         cfg.N_CELLS, float(a_val), float(f_a1), float(da),
                         _f32c(potentials, (cfg.N_CELLS,) * 3),
                         acc.ctypes.data if acc is not None else None, int(cfg.N_CPU))
    return positions, velocities


def advance_time(rho, positions, velocities, fgrid, a, da, cfg: Config, acc=None, phi_out=None):
    """integrate.py:9-13."""
    potentials = potential(rho, fgrid, a, cfg)
    if phi_out is not None:
        phi_out[...] = potentials
    fa1 = f(a + da, [cfg.H0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0])
    return integrate(positions, velocities, a, fa1, da, potentials, cfg, acc=acc)


def cell_keys(positions, cfg: Config):
    """(z_c*Nc + y_c)*Nc + x_c per particle (density.py:19-21,37); int64."""
    npart = positions.shape[1]
    keys = np.empty(npart, dtype=np.int64)
    _lib().pmo_cell_keys(_f32c(positions, (3, npart)), npart, cfg.N_CELLS, keys.ctypes.data)
    return keys


def sort_order(positions, cfg: Config):
    """Stable argsort of the cell keys: the order the CUDA deposit must reproduce bit-exactly."""
    return np.argsort(cell_keys(positions, cfg), kind='stable')


def loop_trip_count(cfg: Config) -> int:
    """pmesh.py:30,56,63 verbatim."""
    da = (cfg.A_END - cfg.A_INIT) / cfg.STEPS
    a_current = cfg.A_INIT
    n = 0
    while a_current < cfg.A_END - da:
        a_current += da
        n += 1
    return n


def loop_cadence(cfg, a_start=None, n_file=1, n_plot=0):
    """pmesh.py:30-34,56-74 restated: which loop iterations write a snapshot / a plot, under which
    file index and at which a_current.  cfg needs N_SAVE_FILES and N_PLOTS besides the Config names;
    n_file = 1 is the state after the initial-conditions snapshot (pmesh.py:46-48)."""
    da = (cfg.A_END - cfg.A_INIT) / cfg.STEPS
    da_save = (cfg.A_END - cfg.A_INIT) / cfg.N_SAVE_FILES
    da_plot = (cfg.A_END - cfg.A_INIT) / cfg.N_PLOTS
    a_current = cfg.A_INIT if a_start is None else a_start
    saves, plots, i = [], [], 0
    while a_current < cfg.A_END - da:
        a_current += da
        if a_current >= cfg.A_INIT + n_file * da_save:
            saves.append((i, n_file, a_current))
            n_file += 1
        if a_current >= cfg.A_INIT + n_plot * da_plot:
            plots.append((i, n_plot))
            n_plot += 1
        i += 1
    return saves, plots


def snapshot_units(a, cfg):
    """save_data.py:10-11: (unit_conv_pos [Mpc], unit_conv_vel [km/s]); cfg needs BOX_SIZE."""
    unit_conv_pos = 7.8 * (cfg.BOX_SIZE / (cfg.N_CELLS / 128)) / 10 ** 3
    unit_conv_vel = 0.781 * cfg.BOX_SIZE * cfg.H0 / (a * cfg.N_CELLS / 128)
    return unit_conv_pos, unit_conv_vel


def step(positions, velocities, fgrid, a_current, da, cfg: Config, mass=None):
    """One body of the loop pmesh.py:56-63.  Returns (rho, positions, velocities)."""
    if mass is None:
        mass = (cfg.N_CELLS / cfg.N_PARTS) ** 3  # pmesh.py:28
    rho = density(positions, mass, cfg)
    advance_time(rho, positions, velocities, fgrid, a_current, da, cfg)
    return rho, positions, velocities


# ---------------------------------------------------------------------------------------------
# Synthetic inputs and the P(k) estimator -- shared by tests and bench.
# ---------------------------------------------------------------------------------------------

def lattice_ic(n_parts: int, n_cells: int, seed: int = 38, jitter: float = 2.0, vel_rms: float = 0.0):
    """IC-like particles: the unperturbed lattice of zeldovich.py:79-83 (row-0 coordinate slowest,
    +0.5 offset) plus a seeded uniform(-jitter, jitter) displacement standing in for
    zeldovich.py:89-91 (whose RNG is unseeded), wrapped with % N_CELLS (zeldovich.py:93)."""
    rs = np.random.RandomState(seed)
    res = n_cells / n_parts
    ax = np.linspace(0, n_cells - res, n_parts) + 0.5
    g = np.meshgrid(ax, ax, ax, indexing='ij')
    pos = np.stack([c.reshape(-1) for c in g]).astype(np.float64)
    pos += rs.uniform(-jitter, jitter, size=pos.shape)
    pos = (pos % n_cells).astype(np.float32)
    vel = (vel_rms * rs.standard_normal(pos.shape)).astype(np.float32)
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel)


def power_spectrum(rho, nbins=None):
    """Spherically binned |rho_k|^2 of the density contrast.  The reference has no estimator
    (SURVEY f3); this one is applied identically to oracle and CUDA outputs."""
    nc = rho.shape[0]
    delta = rho.astype(np.float64) / rho.astype(np.float64).mean() - 1.0
    dk = scipy.fft.rfftn(delta)
    p3 = (dk.real ** 2 + dk.imag ** 2) / float(nc) ** 6
    kz = np.fft.fftfreq(nc) * nc
    kx = np.fft.rfftfreq(nc) * nc
    kk = np.sqrt(kz[:, None, None] ** 2 + kz[None, :, None] ** 2 + kx[None, None, :] ** 2)
    nb = nbins or nc // 2
    edges = np.arange(0.5, nb + 0.5)
    which = np.digitize(kk.ravel(), edges)
    w = np.ones_like(p3)
    w[..., 1:(nc + 1) // 2] = 2.0  # Hermitian half counts twice
    num = np.bincount(which, weights=(p3 * w).ravel(), minlength=nb + 1)
    den = np.bincount(which, weights=w.ravel(), minlength=nb + 1)
    sel = slice(1, nb)
    return 0.5 * (edges[:-1] + edges[1:]), num[sel] / np.maximum(den[sel], 1)
