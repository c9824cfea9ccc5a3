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
caller's arrays again.  The package's own entry points that write into caller tensors through raw
pointers (step, integrate, ResidentParticles.store, rho_out=) bump those counters themselves
(`after_raw_write`), so mixing them with the two calls above is safe.  (A write by ANOTHER library
through a raw pointer cannot be seen; call `forget()` after such a write, or disable the mechanism
with `set_enabled(False)` / PM_DROPIN_RESIDENT=0.)  CUDA tensors only: NumPy arrays carry no version.

Write-back.  By default every `advance_time` ends with pm_particles_store: the caller's tensors hold
the new state in the caller's order when the call returns, exactly the in-place contract of the
reference (src/integrate.py:15-25 updates its arguments).  That un-permute is a 4-byte scatter of
24 bytes per particle -- 1.05 ms of a 3.35 ms step at 256^3 particles, for arrays the reference's
loop never reads between two steps (src/pmesh.py:56-63 only hands them to the next call and, every
N-th step, to save_file).  `set_enabled("lazy")` / PM_DROPIN_RESIDENT=lazy defers it: `advance_time`
returns `ResidentView` handles -- torch.Tensor subclasses over the caller's own storage -- and the
scatter runs the first time anybody USES one of them (any torch function or method that touches
data: .cpu(), indexing, arithmetic, data_ptr(), __cuda_array_interface__, np.asarray ...; pure
metadata such as .shape does not count), or at `sync()`.  `density` / `advance_time` accept the
handles (and the original tensors) and continue from the resident state without any write-back.
The one thing lazy mode cannot see is a READ of the original tensor objects through a name the
caller kept instead of the returned handles (`advance_time(...)` with the result ignored, then
`positions.cpu()`): those bytes are stale until `sync()`.  The reference's loop rebinds
`positions, velocities = advance_time(...)`, so it only ever holds the handles.  The package's own
entry points are covered either way: they call `before_raw_access` on their arguments, which runs
the pending write-back when the session mirrors one of them.
"""
import os
import weakref

import torch

try:
    from . import _runtime as rt
except ImportError:  # flat layout
    import _runtime as rt

_mode = os.environ.get("PM_DROPIN_RESIDENT", "1").strip().lower()
_enabled = _mode != "0"
_lazy = _mode == "lazy"
_session = None          # at most one: the reference's loop drives one particle set
_last_rho = None         # (weakref(rho), version, mean) of the last density() result


def set_enabled(flag):
    """True: resident session, write-back at the end of every advance_time (default).  "lazy": write-back
    deferred until the returned handles are used (module docstring).  False: stateless calls."""
    global _enabled, _lazy
    lazy = isinstance(flag, str) and flag.strip().lower() == "lazy"
    if not lazy:
        sync()
    _enabled = bool(flag)
    _lazy = lazy
    if not _enabled:
        forget()


def enabled() -> bool:
    return _enabled


def lazy() -> bool:
    return _enabled and _lazy


def sync():
    """Bring the caller's tensors up to date with the resident state (no-op unless a lazy write-back
    is pending)."""
    if _session is not None:
        _session.flush()


def forget():
    """Drop the session (frees its plan); the next advance_time starts from the caller's arrays."""
    global _session, _last_rho
    if _session is not None:
        _session.flush()
        _session.close()
    _session = None
    _last_rho = None


def _same_storage(a, b):
    try:
        return a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()
    except Exception:      # a tensor without storage
        return False


def before_raw_access(*tensors):
    """One of the package's C-ABI calls is about to read or write these caller tensors through raw
    pointers.  If the live session mirrors one of them (the same object or another view of its
    storage) and a lazy write-back is pending, the caller's bytes are brought up to date first."""
    s = _session
    if s is None or not s.pending:
        return
    mine = [r() for r in (s.pos_ref, s.vel_ref)]
    for t in tensors:
        t = unwrap(t) if t is not None else None
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            continue
        if any(m is not None and (m is t or _same_storage(m, t)) for m in mine):
            s.flush()
            return


def after_raw_write(*tensors):
    """One of the package's C-ABI calls wrote into these caller tensors through raw pointers, which
    torch's version counters do not see: bump them, as an in-place torch op would have, so that a
    session (or a remembered density mean) that mirrors one of the tensors ends instead of
    continuing from a resident state the caller has just overwritten."""
    for t in tensors:
        t = unwrap(t) if t is not None else None
        if isinstance(t, torch.Tensor):
            try:
                torch._C._increment_version([t])
            except TypeError:                      # signature that takes one tensor
                torch._C._increment_version(t)


def _getter(name):
    return getattr(torch.Tensor, name).__get__


# torch functions that read no tensor data: they do not trigger the deferred write-back
_METADATA = frozenset(
    [_getter(n) for n in ("shape", "device", "dtype", "ndim", "is_cuda", "layout", "requires_grad", "is_leaf",
                          "grad_fn", "_version", "names", "is_sparse", "is_quantized", "is_meta")] +
    [getattr(torch.Tensor, n) for n in ("size", "dim", "numel", "nelement", "stride", "is_contiguous", "storage_offset",
                                        "element_size", "get_device", "is_floating_point", "is_complex", "__len__",
                                        "is_pinned", "is_shared", "_is_view", "type")])


class ResidentView(torch.Tensor):
    """What advance_time returns in lazy mode: the caller's tensor (same storage, same version counter)
    whose bytes may be one or more steps behind the resident state until somebody looks."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        if func not in _METADATA:
            sync()
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **(kwargs or {}))


def unwrap(t):
    """The plain tensor behind a ResidentView (no torch function runs, so nothing is written back)."""
    return t.__dict__["_pm_base"] if type(t) is ResidentView else t


class Session:
    def __init__(self, positions, velocities, n_cells):
        try:
            from .pmesh import ResidentParticles
        except ImportError:
            from pmesh import ResidentParticles
        self.n_cells = int(n_cells)
        self.state = ResidentParticles(positions, velocities)      # own plan; loads the caller's arrays
        self.device = positions.device.index
        self.pending = False          # lazy mode: the caller's tensors are behind the resident state
        self.views = None
        self.remember(positions, velocities)

    def defer(self):
        """Lazy mode, after an advance: no write-back now; hand out the handles."""
        pos, vel = self.pos_ref(), self.vel_ref()
        if self.views is None:
            pv, vv = pos.as_subclass(ResidentView), vel.as_subclass(ResidentView)
            pv.__dict__["_pm_base"], vv.__dict__["_pm_base"] = pos, vel      # the views keep the originals alive
            self.views = (pv, vv)
        self.pending = True
        return self.views

    def flush(self):
        if not self.pending or self.state.plan is None:
            return
        self.pending = False
        pos, vel = self.pos_ref(), self.vel_ref()
        if pos is None or vel is None:
            return                        # nobody can read them any more
        self.state._store(pos, vel)       # the session's own write: the version counters (and the session) stay valid

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
        _session.flush()          # lazy mode: the tensors of the old session get their last state first
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
