"""On-device analysis of the density mesh.

The reference has no power-spectrum estimator, but the acceptance check of the B200 step is
"P(k) within 0.1 % of the reference after the full run", so one is needed; this one reuses the
forward half of the hand-written Poisson transform (pm_power_spectrum in include/pmstep.h) and
never leaves the GPU.  `project` is the plane sum of src/plot_helper.py:65-72."""
try:
    from . import _runtime as rt
except ImportError:
    import _runtime as rt
import torch


def power_spectrum(rho, nbins=None):
    """Spherically binned power spectrum of the density contrast rho/mean - 1.

    rho: float32[Nc, Nc, Nc] CUDA tensor (power-of-two Nc).  Returns (k, P): bin centres 1..nbins-1
    in integer frequency units and P(k) = <|delta_k|^2> / Nc^6 per bin, float64 CUDA tensors."""
    n = rho.shape[0]
    rt.check_dev_f32(rho, (n, n, n), "rho")
    nb = int(nbins or n // 2)
    dev = rho.device.index
    plan = rt.get_plan(n, 1, dev)
    psum = torch.empty(nb, dtype=torch.float64, device=rho.device)
    pcnt = torch.empty_like(psum)
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_power_spectrum(plan.handle, rho.data_ptr(), nb, psum.data_ptr(),
                                            pcnt.data_ptr(), rt.stream_ptr(dev)), "pm_power_spectrum")
    mean = rho.mean(dtype=torch.float64)
    p = psum[1:] / pcnt[1:].clamp_min(1.0) / (mean * mean * float(n) ** 6)
    k = torch.arange(1, nb, dtype=torch.float64, device=rho.device)
    return k, p


def project(rho, n_slices):
    """src/plot_helper.py:65-72: sum of the first n_slices [y, x] planes, float64."""
    return rho[:int(n_slices)].sum(dim=0, dtype=torch.float64)
