"""Gaussian random field of the initial density (reference: src/gaussian_random_field.py:9-123),
generated on the GPU through the C ABI (pm_ic_noise, pm_ic_power_spectrum,
pm_ic_gaussian_random_field of include/pmstep.h).

Same call surface as the reference module: ``gaussian_random_field()`` takes no arguments, reads
N_PARTS, BOX_SIZE, POWER, LCDM_TRANSFER_FUNCTION, RANDOM_SEED, A_INIT and the density parameters
from configure_me and returns the float32 field of shape (N_PARTS,)*3 -- here as a CUDA tensor.
Unlike the reference (seeded NumPy draws inside a parallel numba loop) the noise is a
counter-based Philox stream keyed by RANDOM_SEED: the same seed gives the same field bit for bit."""
import torch

try:
    from . import _runtime as rt
except ImportError:  # flat layout (package directory on sys.path)
    import _runtime as rt


def _workspace(n_parts, device):
    nbytes = int(rt.lib().pm_ic_workspace_bytes(int(n_parts)))
    return torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{device}"), nbytes


def gaussian_random_numbers(seed=None, device=None):
    """Two independent float32 standard-normal fields f1, f2 of shape (N_PARTS,)*3 by the polar
    Box-Muller transform (gaussian_random_field.py:31-63)."""
    cfg = rt.config()
    dev = rt.current_device() if device is None else int(device)
    n = int(cfg.N_PARTS)
    seed = int(cfg.RANDOM_SEED if seed is None else seed)
    f1 = torch.empty((n, n, n), dtype=torch.float32, device=f"cuda:{dev}")
    f2 = torch.empty_like(f1)
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_ic_noise(f1.data_ptr(), f2.data_ptr(), n ** 3, seed, rt.stream_ptr(dev)), "pm_ic_noise")
    return f1, f2


def power_spectrum(device=None):
    """The float64 power-spectrum grid of gaussian_random_field.py:91-123 (k = 0 entry 0)."""
    cfg = rt.config()
    dev = rt.current_device() if device is None else int(device)
    n = int(cfg.N_PARTS)
    prm = rt.ic_params(cfg)
    p = torch.empty((n, n, n), dtype=torch.float64, device=f"cuda:{dev}")
    work, nbytes = _workspace(n, dev)
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_ic_power_spectrum(prm, p.data_ptr(), work.data_ptr(), nbytes, rt.stream_ptr(dev)),
                 "pm_ic_power_spectrum")
    return p


def gaussian_random_field(f1=None, f2=None, device=None):
    """gaussian_random_field.py:9-29.  f1, f2: optional noise fields (CUDA float32, (N_PARTS,)*3);
    by default they are drawn from RANDOM_SEED."""
    cfg = rt.config()
    dev = rt.current_device() if device is None else int(device)
    n = int(cfg.N_PARTS)
    if f1 is None or f2 is None:
        f1, f2 = gaussian_random_numbers(device=dev)
    rt.check_dev_f32(f1, (n, n, n), "f1")
    rt.check_dev_f32(f2, (n, n, n), "f2")
    prm = rt.ic_params(cfg)
    density = torch.empty((n, n, n), dtype=torch.float32, device=f"cuda:{dev}")
    work, nbytes = _workspace(n, dev)
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_ic_gaussian_random_field(prm, f1.data_ptr(), f2.data_ptr(), density.data_ptr(),
                                                      work.data_ptr(), nbytes, rt.stream_ptr(dev)),
                 "pm_ic_gaussian_random_field")
    return density
