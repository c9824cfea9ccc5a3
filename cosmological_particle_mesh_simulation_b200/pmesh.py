"""The per-timestep loop (reference: src/pmesh.py:56-63) as single C-ABI calls.

`step` is one loop body on the caller's CUDA tensors (pm_step, original particle order in and
out), `ResidentParticles` keeps the state inside the plan in cell order between steps (the fast
path: pm_step_resident), `step_host` serves state kept in host memory (pm_step_host, copies
overlapped with compute), and `simulator` is the `while a_current < A_END - da` loop itself with
the reference's predicate kept verbatim.  `run` (= `simulator()` without arguments) is
the reference's whole driver, src/pmesh.py:18-79: initial conditions or restart, the loop, the
snapshot/plot cadences and the status line, with the state resident in HBM throughout."""
try:
    from . import _runtime as rt
    from . import _session
    from .cosmology import f
except ImportError:
    import _runtime as rt
    import _session
    from cosmology import f
import numpy as np
import torch


def _fa1(a, da, cfg):
    return f(a + da, [cfg.H0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0])  # src/integrate.py:12


def step(positions, velocities, a, da, mass=None, rho_out=None):
    """rho = density(pos, mass); pos, vel = advance_time(rho, pos, vel, fgrid, a, da) on CUDA
    tensors, in place.  rho_out (optional float32[Nc,Nc,Nc] CUDA tensor) receives the density."""
    cfg = rt.config()
    n = int(cfg.N_CELLS)
    if mass is None:
        mass = (cfg.N_CELLS / cfg.N_PARTS) ** 3  # src/pmesh.py:28
    rt.check_dev_f32(positions, name="positions")
    rt.check_dev_f32(velocities, tuple(positions.shape), "velocities")
    dev = positions.device.index
    npart = positions.shape[1]
    if rho_out is not None:
        rt.check_dev_f32(rho_out, (n, n, n), "rho_out")
    plan = rt.get_plan(n, npart, dev)
    _session.before_raw_access(positions, velocities)
    with torch.cuda.device(dev):
        rt.check(rt.lib().pm_step(plan.handle, positions.data_ptr(), velocities.data_ptr(), npart,
                                  float(mass), float(a), float(da), float(_fa1(a, da, cfg)),
                                  float(cfg.OMEGA_M0),
                                  rho_out.data_ptr() if rho_out is not None else None,
                                  rt.stream_ptr(dev)), "pm_step")
    _session.after_raw_write(positions, velocities, rho_out)     # ends a drop-in session that mirrors them
    return positions, velocities


def step_host(positions, velocities, a, da, mass=None, rho_out=None, device=None):
    """Same loop body for host-resident state: NumPy arrays or CPU tensors (pinned = fast),
    updated in place.  Returns after the results are back in host memory."""
    cfg = rt.config()
    n = int(cfg.N_CELLS)
    if mass is None:
        mass = (cfg.N_CELLS / cfg.N_PARTS) ** 3
    p = rt.as_host_f32(positions)
    v = rt.as_host_f32(velocities, p.shape)
    r = rt.as_host_f32(rho_out, (n, n, n)) if rho_out is not None else None
    # NumPy arrays that the caller keeps for the run (src/pmesh.py:39-52) are page-locked in place the first time
    # they are seen, so that the copies of this call are asynchronous DMA like those from pinned tensors
    for arr, src in ((p, positions), (v, velocities), (r, rho_out)):
        if isinstance(src, np.ndarray):
            rt.pin_host_array(arr)
    dev = rt.current_device() if device is None else int(device)
    plan = rt.get_plan(n, p.shape[1], dev)
    rt.check(rt.lib().pm_step_host(plan.handle, p.ctypes.data, v.ctypes.data, p.shape[1],
                                   float(mass), float(a), float(da), float(_fa1(a, da, cfg)),
                                   float(cfg.OMEGA_M0), r.ctypes.data if r is not None else None),
             "pm_step_host")
    return positions, velocities


class ResidentParticles:
    """Particle state kept in HBM in cell order across steps (pm_particles_load /
    pm_step_resident / pm_particles_store).  The caller's arrays are only read at construction and
    only written by store(), always in ORIGINAL particle order, so results are comparable
    particle for particle with the reference, which never permutes (src/save_data.py:19-24).
    Every ResidentParticles owns its plan (workspace + state), so stateless calls on the same mesh
    and device -- density(), step(), step_host(), which use the shared cached plan -- cannot disturb it."""

    def __init__(self, positions, velocities):
        cfg = rt.config()
        self.n_cells = int(cfg.N_CELLS)
        rt.check_dev_f32(positions, name="positions")
        rt.check_dev_f32(velocities, tuple(positions.shape), "velocities")
        self.device = positions.device.index
        self.np = positions.shape[1]
        self.plan = rt.Plan(self.n_cells, max(self.np, 1), self.device)
        _session.before_raw_access(positions, velocities)
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_particles_load(self.plan.handle, positions.data_ptr(),
                                                velocities.data_ptr(), self.np,
                                                rt.stream_ptr(self.device)), "pm_particles_load")

    def close(self):
        """Free the plan (workspace and resident state)."""
        if self.plan is not None:
            torch.cuda.synchronize(self.device)
            self.plan.close()
            self.plan = None

    def step(self, a, da, mass=None, rho_out=None):
        cfg = rt.config()
        if mass is None:
            mass = (cfg.N_CELLS / cfg.N_PARTS) ** 3  # src/pmesh.py:28
        if rho_out is not None:
            rt.check_dev_f32(rho_out, (self.n_cells,) * 3, "rho_out")
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_step_resident(
                self.plan.handle, float(mass), float(a), float(da), float(_fa1(a, da, cfg)),
                float(cfg.OMEGA_M0), rho_out.data_ptr() if rho_out is not None else None,
                rt.stream_ptr(self.device)), "pm_step_resident")
        _session.after_raw_write(rho_out)

    def store(self, positions, velocities):
        """Write the state into the caller's (3, Np) CUDA tensors in original particle order."""
        _session.before_raw_access(positions, velocities)
        self._store(positions, velocities)
        _session.after_raw_write(positions, velocities)
        return positions, velocities

    def _store(self, positions, velocities):
        rt.check_dev_f32(positions, (3, self.np), "positions")
        rt.check_dev_f32(velocities, (3, self.np), "velocities")
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_particles_store(self.plan.handle, positions.data_ptr(),
                                                 velocities.data_ptr(), rt.stream_ptr(self.device)),
                     "pm_particles_store")
        return positions, velocities

    def set_sort_mode(self, mode):
        """"auto": re-sort only the particles whose cell changed and merge them into the rest
        (falls back to the full sort by itself); "full": radix-sort every particle every step.
        Both give the same storage order bit for bit (include/pmstep.h, pm_plan_set_sort_mode)."""
        code = {"auto": 0, "full": 1}[mode]
        rt.check(rt.lib().pm_plan_set_sort_mode(self.plan.handle, code), "pm_plan_set_sort_mode")

    def sort_stats(self):
        """(entries, movers, mode) of the last sort; mode is "full" or "incremental"."""
        import ctypes
        n, m, mode = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int(0)
        rt.check(rt.lib().pm_plan_sort_stats(self.plan.handle, ctypes.byref(n), ctypes.byref(m),
                                             ctypes.byref(mode)), "pm_plan_sort_stats")
        return n.value, m.value, {1: "full", 2: "incremental"}.get(mode.value, "none")

    def block_stats(self, rows_per_block=None, cap=None):
        """Skew diagnostics from the row table of the last sort (pm_plan_block_stats): dict with the
        number of row blocks, how many hold more than `cap` particles, the particles beyond `cap` and
        the fullest block.  Defaults: the tile shape of the tiled gather kernel."""
        import ctypes
        rb, cp = ctypes.c_int(0), ctypes.c_int(0)
        rt.check(rt.lib().pm_plan_gather_tile(self.plan.handle, ctypes.byref(rb), ctypes.byref(cp)), "pm_plan_gather_tile")
        rows_per_block = rb.value if rows_per_block is None else int(rows_per_block)
        cap = cp.value if cap is None else int(cap)
        out = (ctypes.c_int64 * 4)()
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_plan_block_stats(self.plan.handle, rows_per_block, cap, out,
                                                  rt.stream_ptr(self.device)), "pm_plan_block_stats")
        return {"rows_per_block": rows_per_block, "cap": cap, "blocks": int(out[0]), "blocks_over_cap": int(out[1]),
                "particles_over_cap": int(out[2]), "max_block_particles": int(out[3])}

    def gather_items(self):
        """(heavy, light, overflow) of the last tiled gather's work list (pm_plan_gather_items)."""
        import ctypes
        h, l, o = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int(0)
        rt.check(rt.lib().pm_plan_gather_items(self.plan.handle, ctypes.byref(h), ctypes.byref(l), ctypes.byref(o)),
                 "pm_plan_gather_items")
        return h.value, l.value, o.value

    def order(self):
        """Original index of the particle in each storage slot (int32 CUDA tensor)."""
        ids = torch.empty(self.np, dtype=torch.int32, device=f"cuda:{self.device}")
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_particles_order(self.plan.handle, ids.data_ptr(),
                                                 rt.stream_ptr(self.device)), "pm_particles_order")
        return ids


def loop_scale_factors(cfg=None):
    """The (a_current, da) pairs the loop of src/pmesh.py:30,56,63 visits, predicate verbatim."""
    cfg = cfg or rt.config()
    da = (cfg.A_END - cfg.A_INIT) / cfg.STEPS
    a_current = cfg.A_INIT
    out = []
    while a_current < cfg.A_END - da:
        out.append((a_current, da))
        a_current += da
    return out


def simulator(positions=None, velocities=None, on_step=None, max_steps=None):
    """Without arguments: the reference's driver, src/pmesh.py:18-79 (see `run`).

    With CUDA tensors: run the loop of src/pmesh.py:56-63 on them.  on_step(i, a_current, rho,
    positions, velocities) is called after each step with the PRE-step density and POST-step
    particles, the pairing the reference's save_file sees (src/pmesh.py:60-67)."""
    if positions is None:
        return run(max_steps=max_steps)
    cfg = rt.config()
    n = int(cfg.N_CELLS)
    rho = torch.empty((n, n, n), dtype=torch.float32, device=positions.device) if on_step else None
    state = ResidentParticles(positions, velocities)
    for i, (a_current, da) in enumerate(loop_scale_factors(cfg)):
        if max_steps is not None and i >= max_steps:
            break
        state.step(a_current, da, rho_out=rho)
        if on_step:
            state.store(positions, velocities)
            on_step(i, a_current + da, rho, positions, velocities)
    return state.store(positions, velocities)


def run(max_steps=None, device=None):
    """The reference's whole driver, src/pmesh.py:18-79, statement for statement: initial
    conditions or restart (:39-52), the integration loop with its predicate (:56-63), the snapshot
    and plot cadences (:65-74) and the status line (:76-77, :81-84).

    B200 specifics: the particle state stays in HBM in cell order for the whole run
    (ResidentParticles); the density mesh is only written out of the step, and the particles only
    brought back to original order, on the steps whose cadence test fires (decided before the step
    from the same floating-point expressions the reference evaluates after it); snapshots leave
    through save_data's copy stream and writer thread, so the following steps overlap the
    device->host copy and the disk write.  Returns (positions, velocities, a_current)."""
    from time import time
    try:
        from .gaussian_random_field import gaussian_random_field
        from .zeldovich import zeldovich
        from .save_data import save_file, from_file, wait
        from .plot_helper import plot_step, plot_grf, plot_projection
    except ImportError:
        from gaussian_random_field import gaussian_random_field
        from zeldovich import zeldovich
        from save_data import save_file, from_file, wait
        from plot_helper import plot_step, plot_grf, plot_projection
    cfg = rt.config()
    dev = rt.current_device() if device is None else int(device)
    N_CELLS, N_PARTS, STEPS = int(cfg.N_CELLS), int(cfg.N_PARTS), cfg.STEPS
    A_INIT, A_END = cfg.A_INIT, cfg.A_END
    flag = lambda name: bool(getattr(cfg, name, False))    # noqa: E731

    print('Starting the simulation for {}^3 particles'.format(N_PARTS), 'with {}^3 grid cells'.format(N_CELLS))
    particle_mass = 1.32 * 10 ** 5 * (cfg.OMEGA_M0 * cfg.H0 ** 2) * (cfg.BOX_SIZE / (N_PARTS / 128)) ** 3
    print('Particle mass in solar mass for this configuration: {:.3E}'.format(particle_mass))
    dens_contrast = (N_CELLS / N_PARTS) ** 3
    print('Particle mass in code units: {:.3e}'.format(dens_contrast))
    da = (A_END - A_INIT) / STEPS
    da_save = (A_END - A_INIT) / cfg.N_SAVE_FILES
    da_plot = (A_END - A_INIT) / cfg.N_PLOTS
    a_current = A_INIT
    n_file = 0
    n_plot = 0

    with torch.cuda.device(dev):
        if flag("RESTART"):
            print("Restarting from file data.{}.hdf5".format(cfg.RESTART_FROM_N))
            n_file = cfg.RESTART_FROM_N
            n_plot = (cfg.RESTART_FROM_N) / cfg.N_SAVE_FILES * cfg.N_PLOTS
            positions, velocities, a_current = from_file(n_file, device=dev)
        else:
            rho = gaussian_random_field(device=dev)
            positions, velocities = zeldovich(rho)
            if flag("SAVE_DATA"):
                save_file(rho, positions, velocities, 0, a_current)
                n_file += 1
            if flag("PLOT_GRF"):
                plot_grf(rho)
            del rho

        rho = torch.empty((N_CELLS,) * 3, dtype=torch.float32, device=positions.device)
        state = ResidentParticles(positions, velocities)       # fourier_grid(): tables live in the plan
        print('Starting the integrations...')
        n_steps = 0
        while a_current < A_END - da:
            if max_steps is not None and n_steps >= max_steps:
                break
            start_time = time()
            a_next = a_current + da                                  # the value the cadence tests see
            saving = a_next >= A_INIT + n_file * da_save
            plotting = a_next >= A_INIT + n_plot * da_plot
            want_save = saving and flag("SAVE_DATA")
            want_plot = plotting and (flag("PLOT_STEPS") or flag("PLOT_PROJECTIONS"))
            need_rho = want_plot or (want_save and flag("SAVE_DENSITY"))

            state.step(a_current, da, mass=dens_contrast, rho_out=rho if need_rho else None)
            a_current += da

            if saving:
                if want_save:
                    state.store(positions, velocities)
                    save_file(rho if need_rho else None, positions, velocities, n_file, a_current)
                n_file += 1
            if plotting:
                if flag("PLOT_STEPS"):
                    plot_step(rho, n_plot)
                if flag("PLOT_PROJECTIONS"):
                    plot_projection(rho, n_plot, 15)
                n_plot += 1
            if flag("PRINT_STATUS"):
                torch.cuda.synchronize(dev)          # the reference's line reports a finished step
                print_status(a_current, start_time, cfg)
            n_steps += 1
        state.store(positions, velocities)
        torch.cuda.synchronize(dev)
    wait()
    return positions, velocities, a_current


def run_slabs(comm=None, max_steps=None, device=None, peers=True):
    """`run` for a slab-decomposed box: the same driver (src/pmesh.py:18-79 -- initial conditions or
    restart, loop predicate, snapshot / plot cadences, status line) with the mesh cut into z slabs
    over the ranks of `comm` (slab.DistComm(): one process per GPU under torchrun, the default;
    slab.LocalComm(P): all ranks in this process on one GPU).

    Initial conditions are generated slab by slab (slab_ic.py: no rank holds the N_PARTS^3 lattice);
    a restart reads the snapshot on every rank and keeps the rank's share.  The FFT transposes,
    ghost planes and migration go through peer memory when it can be set up (`peers`), NCCL
    otherwise.  On a cadence step the particles are collected in original order and the density
    slabs concatenated; the rank that holds slab 0 writes the snapshot and the images, so the
    files are those of the single-GPU run.  Returns (ranks, a_current); the caller closes the ranks
    (slab.release_peers first when peers were set up)."""
    from time import time
    try:
        from . import slab, slab_ic
        from .save_data import save_file, from_file, wait
        from .plot_helper import plot_step, plot_projection
    except ImportError:
        import slab
        import slab_ic
        from save_data import save_file, from_file, wait
        from plot_helper import plot_step, plot_projection
    cfg = rt.config()
    dev = rt.current_device() if device is None else int(device)
    comm = comm if comm is not None else slab.DistComm()
    N_CELLS, N_PARTS = int(cfg.N_CELLS), int(cfg.N_PARTS)
    A_INIT, A_END = cfg.A_INIT, cfg.A_END
    flag = lambda name: bool(getattr(cfg, name, False))    # noqa: E731
    writer = 0 in comm.local_ranks                          # the process that holds slab 0 does the I/O
    npart = N_PARTS ** 3
    dens_contrast = (N_CELLS / N_PARTS) ** 3
    da = (A_END - A_INIT) / cfg.STEPS
    da_save = (A_END - A_INIT) / cfg.N_SAVE_FILES
    da_plot = (A_END - A_INIT) / cfg.N_PLOTS
    a_current = A_INIT
    n_file = 0
    n_plot = 0
    if writer:
        print('Starting the simulation for {}^3 particles'.format(N_PARTS), 'with {}^3 grid cells'.format(N_CELLS),
              'on {} slabs'.format(comm.nranks))

    def gather_slabs(pieces, dim):
        """The whole mesh from the ranks' slabs along `dim` (every process gets it)."""
        if isinstance(comm, slab.LocalComm) or comm.nranks == 1:
            return torch.cat(list(pieces), dim=dim)
        mine = pieces[0].contiguous()
        parts = [torch.empty_like(mine) for _ in range(comm.nranks)]
        comm.dist.all_gather(parts, mine, group=comm.group)
        return torch.cat(parts, dim=dim)

    def full_density(ranks):
        """[Nc, Nc, Nc] density of the last deposit."""
        return gather_slabs([r.buf["RHO"] for r in ranks], 0)

    with torch.cuda.device(dev):
        if flag("RESTART"):
            n_file = cfg.RESTART_FROM_N
            n_plot = (cfg.RESTART_FROM_N) / cfg.N_SAVE_FILES * cfg.N_PLOTS
            positions, velocities, a_current = from_file(n_file, device=dev)
            ranks = slab.make_ranks(N_CELLS, positions, velocities, comm, device=dev)
            del positions, velocities
        else:
            ranks, grf = slab_ic.make_ranks_from_ic(comm, cfg=cfg, device=dev, return_density=True)
            if flag("SAVE_DATA"):
                # the t = 0 snapshot holds the initial Gaussian field (src/pmesh.py:44): assembled from the
                # ranks' columns only here, and only when it is written
                rho0 = gather_slabs(grf, 2) if flag("SAVE_DENSITY") else None
                positions, velocities = slab.collect(ranks, comm, npart)
                if writer:
                    save_file(rho0, positions, velocities, 0, a_current)
                del positions, velocities, rho0
                n_file += 1
            del grf
        if peers and comm.nranks > 1 and slab.setup_peers(ranks, comm):
            slab.setup_ghost_peers(ranks, comm)
        n_steps = 0
        while a_current < A_END - da:
            if max_steps is not None and n_steps >= max_steps:
                break
            start_time = time()
            a_next = a_current + da
            saving = a_next >= A_INIT + n_file * da_save
            plotting = a_next >= A_INIT + n_plot * da_plot
            want_save = saving and flag("SAVE_DATA")
            want_plot = plotting and (flag("PLOT_STEPS") or flag("PLOT_PROJECTIONS"))
            need_rho = want_plot or (want_save and flag("SAVE_DENSITY"))

            slab.slab_step(ranks, comm, a_current, da, mass=dens_contrast, cfg=cfg)
            a_current += da

            rho = full_density(ranks) if need_rho else None          # the PRE-step density
            if saving:
                if want_save:
                    positions, velocities = slab.collect(ranks, comm, npart)
                    if writer:
                        save_file(rho if flag("SAVE_DENSITY") else None, positions, velocities, n_file, a_current)
                    del positions, velocities
                n_file += 1
            if plotting:
                if writer and flag("PLOT_STEPS"):
                    plot_step(rho, n_plot)
                if writer and flag("PLOT_PROJECTIONS"):
                    plot_projection(rho, n_plot, 15)
                n_plot += 1
            del rho
            if writer and flag("PRINT_STATUS"):
                torch.cuda.synchronize(dev)
                print_status(a_current, start_time, cfg)
            n_steps += 1
        torch.cuda.synchronize(dev)
    wait()
    return ranks, a_current


def print_status(a_current, start_time, cfg=None):
    """src/pmesh.py:81-84."""
    from time import time
    cfg = cfg or rt.config()
    percentile = 100 * (a_current - cfg.A_INIT) / (cfg.A_END - cfg.A_INIT)
    print("%.5f" % percentile, "%", " Save step time: --- %.5f seconds ---" % (time() - start_time))


if __name__ == "__main__":      # src/pmesh.py:86-93 (`python pmesh.py` from the package directory)
    import os
    from time import time
    _t0 = time()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:      # under torchrun: one process per GPU, the box cut into slabs
        import torch.distributed as _dist
        _local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(_local)
        _dist.init_process_group("nccl", device_id=torch.device(f"cuda:{_local}"))
        try:
            from . import slab as _slab
        except ImportError:
            import slab as _slab
        _comm = _slab.DistComm()
        _ranks, _ = run_slabs(_comm, device=_local)
        _slab.release_peers(_ranks, _comm)
        for _r in _ranks:
            _r.close()
        _dist.barrier()
        _dist.destroy_process_group()
        if _comm.rank != 0:
            raise SystemExit(0)
    else:
        simulator()
    print("Finished in")
    print("--- %.2f seconds ---" % (time() - _t0))
