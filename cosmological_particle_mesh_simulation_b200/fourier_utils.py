"""fourier_grid() -- the Green's-function table of the Poisson solve
(reference: src/fourier_utils.py:5-16).

The reference returns a float32[Nc, Nc, Nc] array 1/(sin^2(kz/2)+sin^2(ky/2)+sin^2(kx/2)) that
its driver only ever hands back to advance_time (src/pmesh.py:54,61).  The B200 path never reads
such a table (4 B/cell/step of extra HBM traffic; 34 GB at 2048^3): the Green's kernel rebuilds
each factor from an Nc-entry sin^2 table.  fourier_grid() therefore returns an opaque handle
that owns the solver plan; `to_array()` materialises the reference's table (DC entry = 0, which
the reference leaves uninitialised) for inspection and parity checks."""
try:
    from . import _runtime as rt
except ImportError:
    import _runtime as rt
import torch


class FourierGrid:
    def __init__(self, n_cells, device):
        self.n_cells = int(n_cells)
        self.device = int(device)
        self.shape = (self.n_cells,) * 3
        self.dtype = torch.float32

    def plan(self, np_needed=1):
        return rt.get_plan(self.n_cells, np_needed, self.device)

    def to_array(self):
        out = torch.empty(self.shape, dtype=torch.float32, device=f"cuda:{self.device}")
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_fourier_grid(self.plan().handle, out.data_ptr(),
                                              rt.stream_ptr(self.device)), "pm_fourier_grid")
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.to_array().cpu().numpy()
        return a if dtype is None else a.astype(dtype)


def fourier_grid():
    n_cells = int(rt.config().N_CELLS)
    dev = rt.current_device()
    grid = FourierGrid(n_cells, dev)
    grid.plan()
    return grid
