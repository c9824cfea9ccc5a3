"""Host-side scalar cosmology helpers (reference: src/cosmology.py:11-35).  Only `f` is on the
per-step path (src/integrate.py:12); it stays a host double exactly as in the reference."""
import math


def H(a, H0, cosmology):
    """src/cosmology.py:11-18."""
    omega_m0, omega_l0, omega_k0 = cosmology[0], cosmology[1], cosmology[2]
    return math.sqrt(H0 ** 2 * (omega_m0 / a ** 3 + omega_k0 / a ** 2 + omega_l0))


def f(a, cosmology):
    """src/cosmology.py:20-27: reciprocal of the time derivative of a, times H0.

    The loop calls it as ``f(a+da, [H0, OMEGA_LAMBDA0, OMEGA_K0])`` (src/integrate.py:12), i.e.
    with H0 in the Omega_m slot; `integrate.advance_time` reproduces that call verbatim."""
    omegaM, omegaL, omegaK = cosmology[0], cosmology[1], cosmology[2]
    return 1 / math.sqrt((omegaM + omegaK * a + omegaL * a ** 3) / a)


def Dt(a, cosmology):
    """src/cosmology.py:29-35."""
    omegaM, omegaL, omegaK = cosmology[0], cosmology[1], cosmology[2]
    return 5 / 2 / omegaM / (omegaM ** (4 / 7) - omegaL + (1 + omegaM / 2) * (1 + omegaL / 70)) * a
