"""Zel'dovich initial positions and velocities (reference: src/zeldovich.py:10-100) on the GPU
through pm_ic_zeldovich / pm_ic_jitter of include/pmstep.h.

``zeldovich(density)`` has the reference's signature: the float32 Gaussian random field of shape
(N_PARTS,)*3 in, ``(positions, velocities)`` float32 ``[3, N_PARTS**3]`` out, in the reference's
particle order (lattice index (i0*N + i1)*N + i2, direction d displaced along array axis d).  The
per-particle uniform(-2, 2) jitter of zeldovich.py:89-91, which the reference draws from the unseeded
stdlib ``random``, is a Philox stream keyed by RANDOM_SEED here, or an explicit array."""
import torch

try:
    from . import _runtime as rt
    from .gaussian_random_field import _workspace
except ImportError:  # flat layout (package directory on sys.path)
    import _runtime as rt
    from gaussian_random_field import _workspace


def jitter(seed=None, device=None):
    """float64 [3, N_PARTS**3] uniform(-2, 2)."""
    cfg = rt.config()
    dev = rt.current_device() if device is None else int(device)
    n3 = int(cfg.N_PARTS) ** 3
    seed = int(cfg.RANDOM_SEED if seed is None else seed)
    j = torch.empty((3, n3), dtype=torch.float64, device=f"cuda:{dev}")
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_ic_jitter(j.data_ptr(), n3, seed, rt.stream_ptr(dev)), "pm_ic_jitter")
    return j


def zeldovich(density, jitter_field=None):
    """zeldovich.py:10-22."""
    cfg = rt.config()
    n = int(cfg.N_PARTS)
    if rt.is_host(density):
        density = rt.to_device(density, rt.current_device())
    rt.check_dev_f32(density, (n, n, n), "density")
    dev = density.device.index
    if jitter_field is None:
        jitter_field = jitter(device=dev)
    if not (isinstance(jitter_field, torch.Tensor) and jitter_field.is_cuda and jitter_field.dtype == torch.float64
            and tuple(jitter_field.shape) == (3, n ** 3) and jitter_field.is_contiguous()):
        raise TypeError("jitter_field: expected a contiguous CUDA float64 tensor of shape (3, N_PARTS**3)")
    prm = rt.ic_params(cfg)
    positions = torch.empty((3, n ** 3), dtype=torch.float32, device=density.device)
    velocities = torch.empty_like(positions)
    work, nbytes = _workspace(n, dev)
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_ic_zeldovich(prm, density.data_ptr(), jitter_field.data_ptr(), positions.data_ptr(),
                                          velocities.data_ptr(), work.data_ptr(), nbytes, rt.stream_ptr(dev)),
                 "pm_ic_zeldovich")
    return positions, velocities
