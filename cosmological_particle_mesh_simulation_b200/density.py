"""density(positions, mass) -- cloud-in-cell mass deposit (reference: src/density.py:7-48).

B200 path: pm_deposit_cic = cell keys -> stable radix sort by cell -> deterministic
warp-segmented scatter (csrc/pm_particles.cu).  Same signature and result layout as the
reference: positions float32[3, Np] (row 0 = x), returns a fresh float32[Nc, Nc, Nc] indexed
[z, y, x]; N_CELLS comes from configure_me."""
try:
    from . import _runtime as rt
    from . import _session
except ImportError:  # dropped into a flat source tree like the reference's
    import _runtime as rt
    import _session
import torch


def _density_device(positions, mass, n_cells, out=None):
    rt.check_dev_f32(positions, name="positions")
    if positions.dim() != 2 or positions.shape[0] != 3:
        raise ValueError("positions must have shape (3, Np)")
    dev = positions.device.index
    npart = positions.shape[1]
    rho = out if out is not None else torch.empty((n_cells,) * 3, dtype=torch.float32,
                                                  device=positions.device)
    sess = _session.session_for_density(positions, n_cells)
    if sess is not None:
        # these are the positions advance_time wrote last step, untouched: the cell-ordered resident copy
        # is deposited (incremental re-sort) instead of sorting the caller's array from scratch
        sess.deposit(mass, rho)
    else:
        plan = rt.get_plan(n_cells, npart, dev)
        _session.before_raw_access(positions)
        with torch.cuda.device(dev):
            rt.check(rt.lib().pm_deposit_cic(plan.handle, positions.data_ptr(), npart, float(mass),
                                             rho.data_ptr(), rt.stream_ptr(dev)), "pm_deposit_cic")
    _session.note_density(rho, float(mass) * npart / float(n_cells) ** 3)
    return rho


def density(positions, mass):
    n_cells = int(rt.config().N_CELLS)
    positions = _session.unwrap(positions)      # lazy-mode handle of the resident session: no write-back needed
    if rt.is_host(positions):
        dev = rt.current_device()
        rho = _density_device(rt.to_device(positions, dev), mass, n_cells)
        return rho.cpu() if isinstance(positions, torch.Tensor) else rt.to_host_array(rho)
    return _density_device(positions, mass, n_cells)
