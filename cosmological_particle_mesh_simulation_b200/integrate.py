"""advance_time / integrate -- force interpolation and leapfrog kick + drift
(reference: src/integrate.py:9-97).

B200 path: pm_gather_kick_drift, one fused kernel (csrc/pm_particles.cu): CIC weights from the
pre-step position, central-difference force at the 8 corners, kick, drift with periodic wrap.
positions and velocities are updated IN PLACE and returned, like the reference."""
try:
    from . import _runtime as rt
    from . import _session
    from .cosmology import f
    from .potential import _potential_device
    from .fourier_utils import FourierGrid
except ImportError:
    import _runtime as rt
    import _session
    from cosmology import f
    from potential import _potential_device
    from fourier_utils import FourierGrid
import numpy as np
import torch


def _integrate_device(positions, velocities, a_val, f_a1, da, potentials, acc=None):
    n = potentials.shape[0]
    rt.check_dev_f32(potentials, (n, n, n), "potentials")
    rt.check_dev_f32(positions, name="positions")
    rt.check_dev_f32(velocities, tuple(positions.shape), "velocities")
    if positions.dim() != 2 or positions.shape[0] != 3:
        raise ValueError("positions must have shape (3, Np)")
    dev = positions.device.index
    npart = positions.shape[1]
    plan = rt.get_plan(n, 1, dev)
    if acc is not None:
        rt.check_dev_f32(acc, (3, npart), "acc")
    _session.before_raw_access(positions, velocities)
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_gather_kick_drift(
            plan.handle, positions.data_ptr(), velocities.data_ptr(), npart, potentials.data_ptr(),
            float(a_val), float(f_a1), float(da), acc.data_ptr() if acc is not None else None,
            rt.stream_ptr(dev)), "pm_gather_kick_drift")
    _session.after_raw_write(positions, velocities, acc)          # ends a drop-in session that mirrors them
    return positions, velocities


def _copy_back(dst, src_dev):
    rt.copy_to_host(dst, src_dev)      # straight into the caller's array (page-locked in place when it is large)


def integrate(positions, velocities, a_val, f_a1, da, potentials):
    """src/integrate.py:15-25."""
    if rt.is_host(positions):
        dev = rt.current_device()
        p, v = rt.to_device(positions, dev), rt.to_device(velocities, dev)
        phi = rt.to_device(potentials, dev) if rt.is_host(potentials) else potentials
        _integrate_device(p, v, a_val, f_a1, da, phi)
        _copy_back(positions, p)
        _copy_back(velocities, v)
        return positions, velocities
    return _integrate_device(positions, velocities, a_val, f_a1, da, potentials)


def advance_time(density, positions, velocities, fgrid, a, da):
    """src/integrate.py:9-13, including the argument order of the f() call (H0 lands in the
    Omega_m slot of src/cosmology.py:23)."""
    cfg = rt.config()
    fa1 = f(a + da, [cfg.H0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0])
    positions, velocities = _session.unwrap(positions), _session.unwrap(velocities)   # lazy-mode handles
    if rt.is_host(positions):
        dev = rt.current_device()
        rho = rt.to_device(density, dev) if rt.is_host(density) else density
        phi = _potential_device(rho, fgrid, a)
        return integrate(positions, velocities, a, fa1, da, phi)
    if _session.enabled() and isinstance(fgrid, FourierGrid):
        # the resident fast path behind the reference's signature (see _session.py): potential of `density`,
        # gather + kick + drift of the cell-ordered resident copy, result written back in the caller's order
        n = fgrid.n_cells
        rt.check_dev_f32(density, (n, n, n), "density")
        rt.check_dev_f32(positions, name="positions")
        rt.check_dev_f32(velocities, tuple(positions.shape), "velocities")
        if positions.dim() != 2 or positions.shape[0] != 3:
            raise ValueError("positions must have shape (3, Np)")
        sess = _session.session_for_advance(positions, velocities, n)
        sess.advance(density, _session.known_mean(density), a, da, fa1, cfg.OMEGA_M0)
        if _session.lazy():
            return sess.defer()           # handles over the caller's storage; written back on first use
        sess.state._store(positions, velocities)      # the session's own write-back: it stays valid
        return positions, velocities
    phi = _potential_device(density, fgrid, a)
    return _integrate_device(positions, velocities, a, fa1, da, phi)
