"""Runtime glue between the Python drop-in modules and libpmstep.so (include/pmstep.h).

PyTorch is used for device memory, streams and (multi-GPU) torch.distributed only; every
computation of the particle-mesh step happens inside the C-ABI library.  There is NO CPU
fallback: if the library is missing or no CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import importlib
import os
import sys
import threading
import weakref

import numpy as np
import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# PM_LIB lets an experiment point at an alternative build of the same ABI (tuning A/B runs).
LIB_PATH = os.environ.get("PM_LIB") or os.path.join(_PKG_DIR, "libpmstep.so")

_lib = None
_lock = threading.Lock()


class PMStepError(RuntimeError):
    """A libpmstep.so entry point returned non-zero."""

    def __init__(self, code, where=""):
        self.code = int(code)
        msg = lib().pm_error_string(self.code).decode() if _lib is not None else "?"
        super().__init__(f"{where}: libpmstep error {self.code}: {msg}")


class ICParams(ctypes.Structure):
    """pm_ic_params of include/pmstep.h."""
    _fields_ = [("n_parts", ctypes.c_int), ("n_cells", ctypes.c_int), ("box_size", ctypes.c_double),
                ("power", ctypes.c_double), ("lcdm_transfer", ctypes.c_int), ("omega_m0", ctypes.c_double),
                ("omega_b0", ctypes.c_double), ("omega_k0", ctypes.c_double),
                ("omega_lambda0", ctypes.c_double), ("h0", ctypes.c_double), ("a_init", ctypes.c_double)]


def ic_params(cfg=None) -> ICParams:
    """pm_ic_params from the configure_me names the reference's IC modules read
    (src/gaussian_random_field.py:4-5, src/zeldovich.py:7-8)."""
    c = cfg if cfg is not None else config()
    return ICParams(int(c.N_PARTS), int(c.N_CELLS), float(c.BOX_SIZE), float(c.POWER),
                    int(bool(c.LCDM_TRANSFER_FUNCTION)), float(c.OMEGA_M0), float(c.OMEGA_B0),
                    float(c.OMEGA_K0), float(c.OMEGA_LAMBDA0), float(c.H0), float(c.A_INIT))


def lib():
    """Load libpmstep.so (once).  Raises if it has not been built -- run
    ``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C <pkg>/csrc``."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU "
                f"fallback.  Build it with `make -C {os.path.join(_PKG_DIR, 'csrc')}`.")
        L = ctypes.CDLL(LIB_PATH)
        i32, i64, f64, vp, sz = (ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p,
                                 ctypes.c_size_t)
        sig = {
            "pm_version": (ctypes.c_char_p, []),
            "pm_error_string": (ctypes.c_char_p, [i32]),
            "pm_last_cufft_status": (i32, []),
            "pm_launch_count": (ctypes.c_uint64, []),
            "pm_plan_workspace_bytes": (sz, [i32, i64]),
            "pm_plan_create": (i32, [ctypes.POINTER(vp), i32, i64, i32]),
            "pm_plan_destroy": (i32, [vp]),
            "pm_plan_n_cells": (i32, [vp]),
            "pm_plan_np_capacity": (i64, [vp]),
            "pm_plan_set_fft_backend": (i32, [vp, i32]),
            "pm_plan_set_sin2_table": (i32, [vp, vp]),
            "pm_plan_set_poisson_options": (i32, [vp, i32, i32]),
            "pm_plan_poisson_options": (i32, [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]),
            "pm_plan_fft_backend": (i32, [vp]),
            "pm_plan_set_fft_fuse": (i32, [vp, i32, i32]),
            "pm_plan_fft_sync_errors": (i32, [vp]),
            "pm_plan_set_fft_variant": (i32, [vp, i32]),
            "pm_plan_set_gather_tiled": (i32, [vp, i32]),
            "pm_plan_gather_tile": (i32, [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]),
            "pm_plan_block_stats": (i32, [vp, i32, i32, ctypes.POINTER(i64), vp]),
            "pm_plan_set_sort_mode": (i32, [vp, i32]),
            "pm_plan_sort_stats": (i32, [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i32)]),
            "pm_fourier_grid": (i32, [vp, vp, vp]),
            "pm_cell_keys": (i32, [vp, vp, i64, vp, vp]),
            "pm_sort_by_cell": (i32, [vp, vp, i64, vp, vp, vp]),
            "pm_deposit_cic": (i32, [vp, vp, i64, f64, vp, vp]),
            "pm_poisson": (i32, [vp, vp, f64, f64, vp, vp]),
            "pm_gather_kick_drift": (i32, [vp, vp, vp, i64, vp, f64, f64, f64, vp, vp]),
            "pm_power_spectrum": (i32, [vp, vp, i32, vp, vp, vp]),
            "pm_step": (i32, [vp, vp, vp, i64, f64, f64, f64, f64, f64, vp, vp]),
            "pm_step_host": (i32, [vp, vp, vp, i64, f64, f64, f64, f64, f64, vp]),
            "pm_particles_load": (i32, [vp, vp, vp, i64, vp]),
            "pm_step_resident": (i32, [vp, f64, f64, f64, f64, f64, vp, vp]),
            "pm_plan_set_graph": (i32, [vp, i32]),
            "pm_plan_graph_replays": (i32, [vp]),
            "pm_plan_gather_items": (i32, [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i32)]),
            "pm_resident_deposit": (i32, [vp, f64, vp, vp]),
            "pm_resident_advance": (i32, [vp, vp, f64, f64, f64, f64, f64, vp]),
            "pm_particles_store": (i32, [vp, vp, vp, vp]),
            "pm_particles_order": (i32, [vp, vp, vp]),
            "pm_particles_count": (i64, [vp]),
            "pm_plan_create_slab": (i32, [ctypes.POINTER(vp), i32, i64, i32, i32, i32]),
            "pm_slab_buffer": (i32, [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(sz)]),
            "pm_slab_load": (i32, [vp, vp, vp, vp, i64, vp]),
            "pm_slab_count": (i64, [vp]),
            "pm_slab_entries": (i64, [vp]),
            "pm_slab_deposit": (i32, [vp, f64, vp]),
            "pm_slab_ghost_add": (i32, [vp, vp]),
            "pm_slab_fft_rows_forward": (i32, [vp, vp]),
            "pm_slab_set_rho_mean": (i32, [vp, f64]),
            "pm_slab_fft_y_forward": (i32, [vp, i32, i32, vp]),
            "pm_slab_fft_z": (i32, [vp, i32, i32, f64, f64, vp]),
            "pm_slab_fft_y_inverse": (i32, [vp, i32, i32, vp]),
            "pm_slab_fft_rows_inverse": (i32, [vp, vp]),
            "pm_slab_gather": (i32, [vp, f64, f64, f64, vp]),
            "pm_slab_migrate_pack": (i32, [vp, vp, vp]),
            "pm_slab_migrate_unpack": (i32, [vp, i64, i64, vp]),
            "pm_slab_export": (i32, [vp, vp, vp, vp, vp, vp]),
            "pm_slab_peer_export": (i32, [vp, vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
            "pm_slab_peer_import": (i32, [vp, i32, vp, ctypes.c_uint64, ctypes.c_uint64]),
            "pm_slab_peer_set": (i32, [vp, i32, vp, vp]),
            "pm_slab_peer_signal": (i32, [vp, i32, vp]),
            "pm_slab_peer_wait": (i32, [vp, i32, vp]),
            "pm_slab_peer_timeouts": (i32, [vp, ctypes.POINTER(ctypes.c_uint32)]),
            "pm_slab_peer_release": (i32, [vp]),
            "pm_slab_fft_y_forward_local": (i32, [vp, i32, i32, vp]),
            "pm_slab_fft_push": (i32, [vp, i32, i32, vp]),
            "pm_slab_fft_pull": (i32, [vp, i32, i32, vp]),
            "pm_slab_fft_y_inverse_local": (i32, [vp, i32, i32, vp]),
            "pm_slab_peer_ghost_export": (i32, [vp, ctypes.POINTER(ctypes.c_uint64)]),
            "pm_slab_peer_ghost_import": (i32, [vp, i32, ctypes.c_uint64]),
            "pm_slab_peer_ghost_set": (i32, [vp, i32, vp]),
            "pm_slab_ghost_push_rho": (i32, [vp, vp]),
            "pm_slab_ghost_wait_rho": (i32, [vp, vp]),
            "pm_slab_ghost_push_phi": (i32, [vp, vp]),
            "pm_slab_ghost_wait_phi": (i32, [vp, vp]),
            "pm_slab_peer_aux_export": (i32, [vp, ctypes.POINTER(ctypes.c_uint64)]),
            "pm_slab_peer_aux_import": (i32, [vp, i32, ctypes.POINTER(ctypes.c_uint64)]),
            "pm_slab_peer_aux_set": (i32, [vp, i32, vp, vp, vp]),
            "pm_slab_aux_buffers": (i32, [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]),
            "pm_slab_migrate_counts_push": (i32, [vp, vp]),
            "pm_slab_migrate_counts_read": (i32, [vp, vp, vp]),
            "pm_slab_migrate_push": (i32, [vp, vp, vp, vp]),
            "pm_slab_migrate_wait": (i32, [vp, vp]),
            "pm_slab_fft_y_forward_push": (i32, [vp, i32, i32, vp]),
            "pm_slab_fft_y_inverse_pull": (i32, [vp, i32, i32, vp]),
            "pm_ic_workspace_bytes": (sz, [i32]),
            "pm_ic_noise": (i32, [vp, vp, i64, ctypes.c_uint64, vp]),
            "pm_ic_jitter": (i32, [vp, i64, ctypes.c_uint64, vp]),
            "pm_ic_power_spectrum": (i32, [ctypes.POINTER(ICParams), vp, vp, sz, vp]),
            "pm_ic_gaussian_random_field": (i32, [ctypes.POINTER(ICParams), vp, vp, vp, vp, sz, vp]),
            "pm_ic_zeldovich": (i32, [ctypes.POINTER(ICParams), vp, vp, vp, vp, vp, sz, vp]),
            "pm_ic_slab_workspace_bytes": (sz, []),
            "pm_ic_noise_range": (i32, [vp, vp, i64, i64, ctypes.c_uint64, vp]),
            "pm_ic_slab_rho_k": (i32, [ctypes.POINTER(ICParams), vp, vp, i32, i32, vp, vp, sz, vp]),
            "pm_ic_slab_fft": (i32, [vp, i32, i32, i32, i32, i32, vp]),
            "pm_ic_slab_real_f32": (i32, [vp, i64, f64, vp, vp]),
            "pm_ic_slab_from_f32": (i32, [vp, i64, vp, vp]),
            "pm_ic_slab_displacement_k": (i32, [ctypes.POINTER(ICParams), i32, vp, i32, i32, vp, vp]),
            "pm_ic_slab_particles": (i32, [ctypes.POINTER(ICParams), i32, vp, i32, i32, ctypes.c_uint64, vp, vp, vp, vp, vp]),
            "pm_step_host_range": (i32, [i64, i32, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i32)]),
            "pm_host_register": (i32, [vp, sz]),
            "pm_host_unregister": (i32, [vp]),
            "pm_plan_profile_begin": (i32, [vp, i32]),
            "pm_plan_profile_read": (i32, [vp, vp, ctypes.POINTER(i32)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch, fail loudly
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


EXPORTED_SYMBOLS = (
    "pm_version", "pm_error_string", "pm_last_cufft_status", "pm_launch_count",
    "pm_plan_workspace_bytes", "pm_plan_create", "pm_plan_destroy", "pm_plan_n_cells",
    "pm_plan_np_capacity", "pm_fourier_grid", "pm_cell_keys", "pm_sort_by_cell", "pm_deposit_cic",
    "pm_poisson", "pm_gather_kick_drift", "pm_step", "pm_step_host", "pm_plan_profile_begin",
    "pm_plan_profile_read", "pm_particles_load", "pm_step_resident", "pm_resident_deposit", "pm_resident_advance", "pm_plan_set_graph", "pm_plan_graph_replays", "pm_plan_gather_items", "pm_particles_store",
    "pm_particles_order", "pm_particles_count", "pm_plan_set_fft_backend", "pm_plan_fft_backend", "pm_plan_set_sin2_table", "pm_plan_set_poisson_options", "pm_plan_poisson_options",
    "pm_plan_create_slab", "pm_slab_buffer", "pm_slab_load", "pm_slab_count", "pm_slab_entries",
    "pm_slab_deposit", "pm_slab_ghost_add", "pm_slab_fft_rows_forward", "pm_slab_set_rho_mean", "pm_slab_fft_y_forward",
    "pm_slab_fft_z", "pm_slab_fft_y_inverse", "pm_slab_fft_rows_inverse", "pm_slab_gather",
    "pm_slab_migrate_pack", "pm_slab_migrate_unpack", "pm_slab_export",
    "pm_slab_peer_export", "pm_slab_peer_import", "pm_slab_peer_set", "pm_slab_peer_signal", "pm_slab_peer_wait",
    "pm_slab_peer_timeouts", "pm_slab_peer_release", "pm_slab_fft_y_forward_local", "pm_slab_fft_push", "pm_slab_fft_pull",
    "pm_slab_fft_y_inverse_local", "pm_slab_fft_y_forward_push", "pm_slab_fft_y_inverse_pull",
    "pm_slab_peer_ghost_export", "pm_slab_peer_ghost_import", "pm_slab_peer_ghost_set", "pm_slab_ghost_push_rho",
    "pm_slab_ghost_wait_rho", "pm_slab_ghost_push_phi", "pm_slab_ghost_wait_phi", "pm_power_spectrum",
    "pm_slab_peer_aux_export", "pm_slab_peer_aux_import", "pm_slab_peer_aux_set", "pm_slab_aux_buffers",
    "pm_slab_migrate_counts_push", "pm_slab_migrate_counts_read", "pm_slab_migrate_push", "pm_slab_migrate_wait",
    "pm_plan_set_sort_mode", "pm_plan_sort_stats", "pm_plan_set_fft_fuse", "pm_plan_fft_sync_errors",
    "pm_plan_set_fft_variant", "pm_plan_set_gather_tiled", "pm_plan_gather_tile", "pm_plan_block_stats", "pm_ic_workspace_bytes", "pm_ic_noise",
    "pm_ic_jitter", "pm_ic_power_spectrum", "pm_ic_gaussian_random_field", "pm_ic_zeldovich",
    "pm_ic_slab_workspace_bytes", "pm_ic_noise_range", "pm_ic_slab_rho_k", "pm_ic_slab_fft", "pm_ic_slab_real_f32",
    "pm_ic_slab_from_f32", "pm_ic_slab_displacement_k", "pm_ic_slab_particles",
    "pm_host_register", "pm_host_unregister", "pm_step_host_range",
)

STAGE_NAMES = ("keys", "sort", "rows", "deposit", "fft_r2c", "green", "fft_c2r", "gather_kick_drift")


def check(rc, where):
    if rc != 0:
        raise PMStepError(rc, where)


def launch_count() -> int:
    return int(lib().pm_launch_count())


# ---------------------------------------------------------------------------------------------
# configuration: which configure_me is read
# ---------------------------------------------------------------------------------------------
_config_override = None


def set_config(cfg):
    """Use `cfg` (module or any object with the configure_me attribute names) instead of the
    configure_me module.  Pass None to go back to module lookup."""
    global _config_override
    _config_override = cfg


def config():
    """The configure_me the drop-in modules read, at call time: an explicit set_config() object,
    else a top-level module named ``configure_me`` (the reference's own file when these modules
    are dropped into its source tree), else this package's configure_me.py."""
    if _config_override is not None:
        return _config_override
    mod = sys.modules.get("configure_me")
    if mod is None:
        try:
            mod = importlib.import_module("configure_me")
        except ImportError:
            mod = importlib.import_module(__package__ + ".configure_me" if __package__ else "configure_me")
    return mod


# ---------------------------------------------------------------------------------------------
# plans
# ---------------------------------------------------------------------------------------------
def reference_sin2_table(n_cells: int):
    """sin^2(k_i/2) for the Nc wavenumbers of one axis, evaluated with the reference's own NumPy
    expressions (src/fourier_utils.py:8-15: k = float32(2*pi*fftfreq(Nc)), np.sin(k/2)**2 in float32).
    The Green's table of fourier_grid() is 1/((s[i] + s[j]) + s[l]); with this s installed in a plan
    (pm_plan_set_sin2_table) the kernels' G equals the reference's table bit for bit."""
    import numpy as np
    scale = 2 * np.pi
    k = np.array(scale * np.fft.fftfreq(int(n_cells)), dtype="float32")
    return np.ascontiguousarray(np.sin(k / 2) ** 2, dtype=np.float32)


def install_reference_tables(handle, n_cells: int):
    s2 = reference_sin2_table(n_cells)
    check(lib().pm_plan_set_sin2_table(handle, s2.ctypes.data), "pm_plan_set_sin2_table")


# Options of the Poisson solve the reference does not have (include/pmstep.h, pm_plan_set_poisson_options):
# (deconvolve, kspace_gradient); (0, 0) = the reference's scheme.  Applied to every single-GPU plan made
# after set_poisson_options() and to the cached ones.
_poisson_options = (0, 0)


def set_poisson_options(deconvolve=0, kspace_gradient=False):
    """deconvolve: 0, 1 or 2 = power of the CIC window divided out of phi_k; kspace_gradient: forces from
    -i k phi_k on three force meshes instead of central differences of phi.  Defaults = the reference."""
    global _poisson_options
    _poisson_options = (int(deconvolve), 1 if kspace_gradient else 0)
    try:
        from . import _session
    except ImportError:
        import _session
    _session.forget()         # a live drop-in session owns a plan made under the old options
    for plan in _plans.values():
        if plan.handle is not None:
            check(lib().pm_plan_set_poisson_options(plan.handle, *_poisson_options), "pm_plan_set_poisson_options")


def poisson_options():
    return _poisson_options


class Plan:
    """Owns one pm_plan* (cuFFT plans, Green's table, all per-step scratch) on one device."""

    def __init__(self, n_cells: int, np_capacity: int, device: int):
        self.n_cells, self.np_capacity, self.device = int(n_cells), int(np_capacity), int(device)
        h = ctypes.c_void_p()
        check(lib().pm_plan_create(ctypes.byref(h), self.n_cells, self.np_capacity, self.device),
              f"pm_plan_create(n_cells={n_cells}, np={np_capacity}, device={device})")
        self.handle = h
        install_reference_tables(h, self.n_cells)
        if _poisson_options != (0, 0):
            check(lib().pm_plan_set_poisson_options(h, *_poisson_options), "pm_plan_set_poisson_options")

    def close(self):
        if getattr(self, "handle", None):
            lib().pm_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plans: dict = {}


def get_plan(n_cells: int, np_needed: int, device: int) -> Plan:
    """Cached plan for (n_cells, device) whose particle capacity covers np_needed."""
    key = (int(n_cells), int(device))
    plan = _plans.get(key)
    if plan is None or plan.np_capacity < np_needed or plan.handle is None:
        if plan is not None:
            torch.cuda.synchronize(device)
            plan.close()
        plan = Plan(n_cells, max(int(np_needed), 1), device)
        _plans[key] = plan
    return plan


def release_plans():
    try:
        from . import _session
    except ImportError:
        import _session
    _session.forget()
    for plan in _plans.values():
        plan.close()
    _plans.clear()


# ---------------------------------------------------------------------------------------------
# tensors
# ---------------------------------------------------------------------------------------------
def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the B200 particle-mesh step has no CPU fallback")


def current_device() -> int:
    require_cuda()
    return torch.cuda.current_device()


def stream_ptr(device: int) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def is_host(x) -> bool:
    return isinstance(x, np.ndarray) or (isinstance(x, torch.Tensor) and not x.is_cuda)


def as_host_f32(x, shape=None):
    """A C-contiguous float32 NumPy view/copy of a NumPy array or CPU tensor."""
    a = x.numpy() if isinstance(x, torch.Tensor) else x
    if a.dtype != np.float32 or not a.flags.c_contiguous:
        raise TypeError("expected a C-contiguous float32 array (reference dtypes, SURVEY Q13)")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def check_dev_f32(t: torch.Tensor, shape=None, name="tensor"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{name}: expected a CUDA tensor")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError(f"{name}: expected contiguous float32")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


# ---------------------------------------------------------------------------------------------
# host arrays of the drop-in calls (NumPy in -> NumPy out, like the reference)
# ---------------------------------------------------------------------------------------------
# The reference's driver keeps positions and velocities in two NumPy arrays for the whole run and hands
# them to density() / advance_time() every step (src/pmesh.py:56-63).  Copied as pageable memory, with a
# fresh host array for every result, that loop spent 0.57 s per step at 256^3 particles on a 512^3 mesh,
# almost all of it in staged copies and first-touch page faults.  So: a large array that keeps coming back
# is page-locked in place once (pm_host_register; un-registered when the array dies), after which torch
# sees it as pinned memory and copies by DMA; results that the reference returns as fresh arrays (the
# density mesh) come from a small pool of pinned buffers that are handed out again once the previous
# array -- and every view of it -- is gone; in-place results are copied straight into the caller's array.
# PM_PIN_HOST_ARRAYS=0 turns both off.
_PIN_MIN_BYTES = 1 << 20
_pin_enabled = os.environ.get("PM_PIN_HOST_ARRAYS", "1").strip() != "0"
_pinned_ranges = {}        # (address, bytes) -> weakref.finalize of the array that was registered
_pin_refused = set()       # ranges the driver would not register (kept so that we ask only once)
_host_pool = []            # [pinned tensor, weakref to the holder of the array handed out]
_HOST_POOL_MAX = 4


def _unpin(key):
    _pinned_ranges.pop(key, None)
    if _lib is not None:
        _lib.pm_host_unregister(key[0])


def pin_host_array(a) -> bool:
    """Page-lock the memory of NumPy array `a` in place (once; released when `a` is garbage-collected)."""
    if not _pin_enabled or not isinstance(a, np.ndarray) or a.nbytes < _PIN_MIN_BYTES or not a.flags.c_contiguous:
        return False
    key = (int(a.ctypes.data), int(a.nbytes))
    if key in _pinned_ranges:
        return True
    if key in _pin_refused:
        return False
    if lib().pm_host_register(key[0], key[1]) != 0:
        _pin_refused.add(key)
        return False
    try:
        fin = weakref.finalize(a, _unpin, key)
        fin.atexit = False         # at interpreter exit the driver releases everything itself
        _pinned_ranges[key] = fin
    except TypeError:          # an ndarray subclass without weak references
        lib().pm_host_unregister(key[0])
        _pin_refused.add(key)
        return False
    return True


class _PinnedHolder:
    """Owner of a pooled pinned buffer as NumPy sees it: every array (and view) made from it keeps it alive."""

    def __init__(self, tensor):
        self._tensor = tensor
        self.__array_interface__ = {"shape": tuple(tensor.shape), "typestr": "<f4", "data": (tensor.data_ptr(), False),
                                    "version": 3, "strides": None}


def to_host_array(t: torch.Tensor):
    """A fresh float32 NumPy array with the contents of CUDA tensor `t`, from the pinned pool when possible."""
    if not _pin_enabled or t.dtype != torch.float32 or t.numel() * 4 < _PIN_MIN_BYTES:
        return t.cpu().numpy()
    slot = None
    for entry in _host_pool:
        if entry[1]() is None and tuple(entry[0].shape) == tuple(t.shape):
            slot = entry
            break
    if slot is None:
        if len(_host_pool) >= _HOST_POOL_MAX:
            free = [e for e in _host_pool if e[1]() is None]
            if not free:
                return t.cpu().numpy()
            _host_pool.remove(free[0])
        try:
            slot = [torch.empty(tuple(t.shape), dtype=torch.float32, pin_memory=True), lambda: None]
        except RuntimeError:
            return t.cpu().numpy()
        _host_pool.append(slot)
    slot[0].copy_(t)                       # synchronous device-to-host DMA
    holder = _PinnedHolder(slot[0])
    slot[1] = weakref.ref(holder)
    return np.asarray(holder)


def copy_to_host(dst, src_dev: torch.Tensor):
    """src_dev (CUDA) into the caller's NumPy array or CPU tensor `dst`, without an intermediate host array."""
    if isinstance(dst, torch.Tensor):
        dst.copy_(src_dev)
        return
    pin_host_array(dst)
    torch.from_numpy(dst).copy_(src_dev)


def to_device(x, device: int) -> torch.Tensor:
    a = as_host_f32(x)
    if isinstance(x, np.ndarray):      # the caller's own array object (a CPU tensor can be pinned by its owner)
        pin_host_array(a)
    return torch.from_numpy(a).to(f"cuda:{device}", non_blocking=False)
