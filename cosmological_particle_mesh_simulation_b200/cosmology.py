"""Host-side scalar cosmology helpers (reference: src/cosmology.py:11-35).  Only `f` is on the
per-step path (src/integrate.py:12); it stays a host scalar exactly as in the reference.

np.sqrt, not math.sqrt: on the restart path save_data.from_file hands back `a` as np.float32, and
under NumPy 2 the reference then evaluates f(a+da) in float32 (Python scalars are weak).  np.sqrt keeps
the dtype of its argument, so a restarted run gets the reference's f_a1 bit for bit; math.sqrt would
promote to float64.  `from cosmology import *` in the reference also exports `np` and the `cosmo`
record (src/cosmology.py:1-9), so both exist here."""
import numpy as np

try:
    from . import _runtime as _rt
except ImportError:  # flat layout (drop-in by module name)
    import _runtime as _rt


class cosmo:
    """src/cosmology.py:5-9: growth factor, Hubble rate and f at A_INIT for the active configuration."""

    def __init__(self):
        c = _rt.config()
        self.Dt = Dt(c.A_INIT, [c.OMEGA_M0, c.OMEGA_LAMBDA0, c.OMEGA_K0])
        self.H01 = H(c.A_INIT, c.H0, [c.OMEGA_M0, c.OMEGA_LAMBDA0, c.OMEGA_K0])
        self.f0 = f(c.A_INIT, [c.OMEGA_M0, c.OMEGA_LAMBDA0, c.OMEGA_K0])


def H(a, H0, cosmology):
    """src/cosmology.py:11-18."""
    omega_m0, omega_l0, omega_k0 = cosmology[0], cosmology[1], cosmology[2]
    return np.sqrt(H0 ** 2 * (omega_m0 / a ** 3 + omega_k0 / a ** 2 + omega_l0))


def f(a, cosmology):
    """src/cosmology.py:20-27: reciprocal of the time derivative of a, times H0.

    The loop calls it as ``f(a+da, [H0, OMEGA_LAMBDA0, OMEGA_K0])`` (src/integrate.py:12), i.e.
    with H0 in the Omega_m slot; `integrate.advance_time` reproduces that call verbatim."""
    omegaM, omegaL, omegaK = cosmology[0], cosmology[1], cosmology[2]
    return 1 / np.sqrt((omegaM + omegaK * a + omegaL * a ** 3) / a)


def Dt(a, cosmology):
    """src/cosmology.py:29-35."""
    omegaM, omegaL, omegaK = cosmology[0], cosmology[1], cosmology[2]
    return 5 / 2 / omegaM / (omegaM ** (4 / 7) - omegaL + (1 + omegaM / 2) * (1 + omegaL / 70)) * a
