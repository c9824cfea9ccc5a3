"""potential(density, fgrid, a) -- FFT Poisson solve (reference: src/potential.py:7-29).

B200 path: pm_poisson = cuFFT R2C -> one fused pass over the half spectrum applying
-3*Omega_m/(8a) * G(k) / Nc^3 with the DC mode zeroed -> cuFFT C2R, all float32
(csrc/pm_poisson.cu).  The reference transforms the real density as complex128 c2c; the results
agree to ~1e-6 relative L2 (tests/test_gpu_parity.py)."""
try:
    from . import _runtime as rt
    from .fourier_utils import FourierGrid
except ImportError:
    import _runtime as rt
    from fourier_utils import FourierGrid
import torch


def _potential_device(rho, fgrid, a, out=None):
    if not isinstance(fgrid, FourierGrid):
        raise TypeError("fgrid must be the handle returned by this package's fourier_grid()")
    n = fgrid.n_cells
    rt.check_dev_f32(rho, (n, n, n), "density")
    dev = rho.device.index
    plan = rt.get_plan(n, 1, dev)
    phi = out if out is not None else torch.empty_like(rho)
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_poisson(plan.handle, rho.data_ptr(), float(a),
                                     float(rt.config().OMEGA_M0), phi.data_ptr(),
                                     rt.stream_ptr(dev)), "pm_poisson")
    return phi


def potential(density, fgrid, a):
    if rt.is_host(density):
        dev = rt.current_device()
        phi = _potential_device(rt.to_device(density, dev), fgrid, a).cpu()
        return phi if isinstance(density, torch.Tensor) else phi.numpy()
    return _potential_device(density, fgrid, a)
