"""CPU oracle for the initial conditions -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/gaussian_random_field.py:9-123 and src/zeldovich.py:10-100 in NumPy
(float64 / complex128, like the reference) with the two things the reference leaves to chance made
explicit inputs:

  * the Gaussian noise fields f1, f2 (the reference seeds NumPy inside a parallel numba function,
    gaussian_random_field.py:31-35: not reproducible);
  * the per-particle uniform(-2, 2) jitter (zeldovich.py:89-91 draws it from the unseeded stdlib
    `random`).

Entries the reference leaves UNINITIALISED (np.power / np.divide with `where=` and no `out=`, at the
k = 0 mode: gaussian_random_field.py:71,109 and zeldovich.py:38) are pinned to 0 here.

Parity status: PINNED for power_spectrum, potential_k, displacement_field_k, zeldovich_positions
and zeldovich_velocities (tests/golden/ic16.npz holds the outputs of the reference's own functions,
produced by oracle/make_golden.py; tests/test_oracle_golden.py compares).  UNPINNED for the two
pyFFTW call sites (gaussian_random_field.py:27-29, zeldovich.py:49-53): they build their arrays with
dtype='cfloat', which NumPy >= 2 rejects, so those lines cannot execute in this image; they are
normalised inverse c2c DFTs (pyFFTW default), restated with scipy.fft.ifftn.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.fft


@dataclass
class ICConfig:
    """configure_me.py names the IC generator reads (configure_me.py:7-12,31-40)."""
    N_PARTS: int = 256
    N_CELLS: int = 512
    BOX_SIZE: float = 100
    N_CPU: int = 1
    POWER: float = 1.00
    LCDM_TRANSFER_FUNCTION: bool = True
    OMEGA_M0: float = 0.31
    OMEGA_B0: float = 0.04
    OMEGA_K0: float = 0.00
    OMEGA_LAMBDA0: float = 0.69
    H0: float = 0.68
    A_INIT: float = 0.01


def Dt(a, cosmology):
    """cosmology.py:29-35."""
    omegaM, omegaL = cosmology[0], cosmology[1]
    return 5 / 2 / omegaM / (omegaM ** (4 / 7) - omegaL + (1 + omegaM / 2) * (1 + omegaL / 70)) * a


def H(a, H0, cosmology):
    """cosmology.py:11-18."""
    omega_m0, omega_l0, omega_k0 = cosmology
    return np.sqrt(H0 ** 2 * (omega_m0 / a ** 3 + omega_k0 / a ** 2 + omega_l0))


def f(a, cosmology):
    """cosmology.py:20-27."""
    omegaM, omegaL, omegaK = cosmology
    return 1 / np.sqrt((omegaM + omegaK * a + omegaL * a ** 3) / a)


def fourier_grid(cfg: ICConfig):
    """gaussian_random_field.py:76-89: |k| on the N_PARTS^3 grid, k = 2 pi N/BOX * fftfreq(N)."""
    n = cfg.N_PARTS
    scale = 2 * np.pi * n / cfg.BOX_SIZE
    ax = scale * np.fft.fftfreq(n)
    lz, ly, lx = np.meshgrid(ax, ax, ax, indexing='ij')
    return np.sqrt(lx ** 2 + ly ** 2 + lz ** 2)


def lcdm_transfer_function(k_grid, cfg: ICConfig):
    """gaussian_random_field.py:65-74 (q = 0 entry of factor2 pinned to 0)."""
    Gamma = cfg.OMEGA_M0 * cfg.H0 * np.exp(-cfg.OMEGA_B0 - cfg.OMEGA_B0 / cfg.OMEGA_M0)
    q = k_grid / Gamma
    factor1 = np.sqrt(1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4)
    factor2 = np.divide(np.log(1 + 2.34 * q) ** 2, (2.34 * q) ** 2, out=np.zeros_like(q), where=q != 0)
    return factor2 / factor1


def power_spectrum(cfg: ICConfig):
    """gaussian_random_field.py:91-123."""
    k_grid = fourier_grid(cfg)
    lcdm = lcdm_transfer_function(k_grid, cfg) if cfg.LCDM_TRANSFER_FUNCTION else 0.
    Npix = cfg.N_PARTS ** 3
    sigma2fluxt = 64 * cfg.H0 ** 2
    nz = k_grid != 0
    if cfg.POWER >= 0.:
        kp = np.power(k_grid, cfg.POWER, out=np.zeros_like(k_grid), where=nz)
        if cfg.LCDM_TRANSFER_FUNCTION:
            summ = np.sum(kp * lcdm)
            A = sigma2fluxt * Npix ** 2 / summ
            p = A * (k_grid) ** cfg.POWER * lcdm
        else:
            summ = np.sum(kp)
            A = sigma2fluxt * Npix ** 2 / summ
            p = A * (k_grid) ** cfg.POWER
    else:
        kgrid_inverse = np.power(k_grid, -cfg.POWER, out=np.zeros_like(k_grid), where=nz)
        div_kgrid = np.divide(1, kgrid_inverse, out=np.zeros_like(k_grid), where=kgrid_inverse != 0)
        summ = np.sum(div_kgrid)
        A = sigma2fluxt * Npix ** 2 / summ
        p = A * div_kgrid
    return p


def gaussian_random_field(f1, f2, cfg: ICConfig):
    """gaussian_random_field.py:9-29 with the noise fields as inputs (float32[N,N,N] each)."""
    D = Dt(cfg.A_INIT, [cfg.OMEGA_M0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0])
    p = power_spectrum(cfg)
    amp = np.sqrt(p * D ** 2)
    rho_k = amp * f1 + 1j * (amp * f2)
    return (scipy.fft.ifftn(rho_k, axes=(0, 1, 2), workers=cfg.N_CPU).real).astype('float32')


def _axes(cfg: ICConfig):
    n = cfg.N_PARTS
    return 2 * np.pi * n / cfg.BOX_SIZE * np.fft.fftfreq(n)


def potential_k(density_k, cfg: ICConfig):
    """zeldovich.py:24-38 (k = 0 entry pinned to 0)."""
    ax = _axes(cfg)
    lz, ly, lx = np.meshgrid(ax, ax, ax, indexing='ij')
    del_sq = -(lx ** 2 + ly ** 2 + lz ** 2)
    return np.divide(density_k, del_sq, out=np.zeros_like(density_k), where=del_sq != 0)


def displacement_field_k(pot_k, direction, cfg: ICConfig):
    """zeldovich.py:56-69: -i * l_direction * phi_k * (N_CELLS/N_PARTS), l along array axis `direction`."""
    resolution = cfg.N_CELLS / cfg.N_PARTS
    ax = _axes(cfg)
    l_direction = np.meshgrid(ax, ax, ax, indexing='ij')[direction]
    return -1.j * l_direction * pot_k * resolution


def displacement_field_one_direction(pot_k, direction, cfg: ICConfig):
    """zeldovich.py:45-54."""
    force_resolution = cfg.N_CELLS / cfg.BOX_SIZE
    df_k = displacement_field_k(pot_k, direction, cfg).astype(np.complex128)
    out = scipy.fft.ifftn(df_k, axes=(0, 1, 2), workers=cfg.N_CPU)
    return np.reshape(out, (cfg.N_PARTS ** 3)).real * force_resolution


def zeldovich_positions(displacement_field, direction, jitter, cfg: ICConfig):
    """zeldovich.py:71-93 with the uniform(-2, 2) jitter of :89-91 as an input (float64[N^3])."""
    D = Dt(cfg.A_INIT, [cfg.OMEGA_M0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0])
    n = cfg.N_PARTS
    mass_resolution = cfg.N_CELLS / n
    sp = np.linspace(0, cfg.N_CELLS - mass_resolution, n) + 0.5
    positions = np.reshape(np.meshgrid(sp, sp, sp, indexing='ij')[direction], n * n * n)
    positions = positions + D * displacement_field
    positions = positions + jitter
    return positions % cfg.N_CELLS


def zeldovich_velocities(displacement_field, cfg: ICConfig):
    """zeldovich.py:95-100."""
    cosmo = [cfg.OMEGA_M0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0]
    dt_0 = Dt(cfg.A_INIT, cosmo)
    h_0 = H(cfg.A_INIT, cfg.H0, cosmo)
    f_0 = f(cfg.A_INIT, cosmo)
    return cfg.A_INIT * f_0 * h_0 * dt_0 * displacement_field


def zeldovich(density, jitter, cfg: ICConfig):
    """zeldovich.py:10-22.  density: float32[N,N,N]; jitter: float64[3, N^3].
    Returns float32 positions[3, N^3], velocities[3, N^3]."""
    n3 = cfg.N_PARTS ** 3
    positions = np.zeros((3, n3), dtype=np.float32)
    velocities = np.zeros((3, n3), dtype=np.float32)
    density_k = np.fft.fftn(density.astype(np.float64))   # complex128, as np.fft did under the reference's NumPy 1.x
    for direction in (0, 1, 2):
        pot_k = potential_k(density_k, cfg)
        disp = displacement_field_one_direction(pot_k, direction, cfg)
        positions[direction, :] = zeldovich_positions(disp, direction, jitter[direction], cfg)
        velocities[direction] = zeldovich_velocities(disp, cfg)
    return positions, velocities
