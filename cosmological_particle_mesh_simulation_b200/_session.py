"""Resident state behind the reference's own call signatures.

The reference's loop body (src/pmesh.py:60-61) is two calls on arrays the caller owns:

    rho = density(positions, dens_contrast)
    positions, velocities = advance_time(rho, positions, velocities, ksq_inverse, a_current, da)

Served statelessly, each call must treat the arrays as new: cell keys, a full sort, a gather in
lattice order -- about 8 ms per step at 256^3 particles on 512^3 cells.  The fast path of this
package (pm_step_resident, 2 ms) keeps the particles inside a plan in cell order between steps.
This module gives the two calls above that speed without changing them: `advance_time` leaves a
*session* behind -- a private plan holding the state it just wrote into the caller's tensors,
remembered together with the identity and the version counters of those tensors.  The next
`density(positions, ...)` / `advance_time(...)` on the SAME, UNMODIFIED tensors continue from the
resident state (pm_resident_deposit / pm_resident_advance) and only write the result back in the
caller's original particle order (pm_particles_store).  Any in-place change the caller makes through
torch bumps the tensor's version counter, which ends the session: the next call starts from the
caller's arrays again.  (A write that bypasses torch -- a raw pointer handed to another library --
cannot be seen; call `forget()` after such a write, or disable the mechanism with
`set_enabled(False)` / PM_DROPIN_RESIDENT=0.)  CUDA tensors only: NumPy arrays carry no version.
"""
import os
import weakref

import torch

try:
    from . import _runtime as rt
except ImportError:  # flat layout
    import _runtime as rt

_enabled = os.environ.get("PM_DROPIN_RESIDENT", "1") != "0"
_session = None          # at most one: the reference's loop drives one particle set
_last_rho = None         # (weakref(rho), version, mean) of the last density() result


def set_enabled(flag: bool):
    global _enabled
    _enabled = bool(flag)
    if not _enabled:
        forget()


def enabled() -> bool:
    return _enabled


def forget():
    """Drop the session (frees its plan); the next advance_time starts from the caller's arrays."""
    global _session, _last_rho
    if _session is not None:
        _session.close()
    _session = None
    _last_rho = None


class Session:
    def __init__(self, positions, velocities, n_cells):
        try:
            from .pmesh import ResidentParticles
        except ImportError:
            from pmesh import ResidentParticles
        self.n_cells = int(n_cells)
        self.state = ResidentParticles(positions, velocities)      # own plan; loads the caller's arrays
        self.device = positions.device.index
        self.remember(positions, velocities)

    def remember(self, positions, velocities):
        self.pos_ref, self.vel_ref = weakref.ref(positions), weakref.ref(velocities)
        self.pos_ver, self.vel_ver = positions._version, velocities._version
        self.pos_ptr, self.vel_ptr = positions.data_ptr(), velocities.data_ptr()

    def matches_positions(self, positions, n_cells):
        return (self.state.plan is not None and self.pos_ref() is positions and positions._version == self.pos_ver
                and positions.data_ptr() == self.pos_ptr and self.n_cells == int(n_cells)
                and positions.device.index == self.device)

    def matches(self, positions, velocities, n_cells):
        return (self.matches_positions(positions, n_cells) and self.vel_ref() is velocities
                and velocities._version == self.vel_ver and velocities.data_ptr() == self.vel_ptr)

    def deposit(self, mass, rho):
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_resident_deposit(self.state.plan.handle, float(mass), rho.data_ptr(),
                                                  rt.stream_ptr(self.device)), "pm_resident_deposit")

    def advance(self, rho, rho_mean, a, da, f_a1, omega_m0):
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_resident_advance(self.state.plan.handle, rho.data_ptr(), float(rho_mean), float(a),
                                                  float(da), float(f_a1), float(omega_m0),
                                                  rt.stream_ptr(self.device)), "pm_resident_advance")

    def close(self):
        self.state.close()


def session_for_density(positions, n_cells):
    """The live session if `positions` is the tensor it last wrote, untouched since; else None."""
    if not _enabled or _session is None:
        return None
    return _session if _session.matches_positions(positions, n_cells) else None


def session_for_advance(positions, velocities, n_cells):
    """The live session for these tensors, or a new one loaded from them."""
    global _session
    if not _enabled:
        return None
    if _session is not None and _session.matches(positions, velocities, n_cells):
        return _session
    if _session is not None:
        _session.close()
        _session = None
    _session = Session(positions, velocities, n_cells)
    return _session


def note_density(rho, mean):
    """density() tells what it returned, so that advance_time need not measure the mean of a mesh
    this library deposited itself (total mass / Nc^3)."""
    global _last_rho
    _last_rho = (weakref.ref(rho), rho._version, rho.data_ptr(), float(mean))


def known_mean(rho):
    if _last_rho is None:
        return float("nan")
    ref, ver, ptr, mean = _last_rho
    return mean if (ref() is rho and rho._version == ver and rho.data_ptr() == ptr) else float("nan")
