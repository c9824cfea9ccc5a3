"""potential(density, fgrid, a) -- FFT Poisson solve (reference: src/potential.py:7-29).

B200 path: pm_poisson.  Power-of-two meshes (32..2048): the hand-written five-pass real transform of
csrc/pm_fft*.cu(h) -- rows and y forward, ONE fused z pass (forward z, times -3*Omega_m/(8a) * G(k) / Nc^3
with the DC mode zeroed, inverse z), y and rows inverse -- all float32, the density mean subtracted on the
way in.  Other mesh sizes, and the deconvolution / spectral-gradient options: cuFFT R2C/C2R around one
fused pass over the half spectrum (csrc/pm_poisson.cu).  G(k) is the reference's table bit for bit
(_runtime.reference_sin2_table).  The reference transforms the real density as complex128 c2c; the
results agree to ~1e-6 relative L2 (tests/test_gpu_parity.py); `pm_plan_set_fft_backend(plan, 2)`
switches to float64 transforms for diagnosis."""
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
        phi = _potential_device(rt.to_device(density, dev), fgrid, a)
        return phi.cpu() if isinstance(density, torch.Tensor) else rt.to_host_array(phi)
    return _potential_device(density, fgrid, a)
