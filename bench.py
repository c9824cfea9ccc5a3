#!/usr/bin/env python
"""bench.py -- particle-steps/s of the full particle-mesh step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one body of the reference loop src/pmesh.py:60-61 (CIC deposit + FFT Poisson solve
+ force gather/kick/drift) over all particles.  Workload at N=1: BASELINE.json configs[1],
256^3 particles on a 512^3 mesh (the README benchmark config), IC-like synthetic particles
(lattice + seeded uniform(-2,2) jitter).

Our arm prints ONE JSON line with
  value      particle-steps/s, state resident in HBM, K steps timed with CUDA events on the
             launching stream (max over ranks);
  e2e        the same metric through the host-buffer C-ABI call pm_step_host(): every step
             uploads positions+velocities from pinned host memory and downloads the results;
  roofline   the dominant hand-written kernel, timed live with CUDA events inside the timed
             region (pm_plan_profile_*), against MEASURED_PEAKS.json hbm_gbs;
  roofline_step  the whole step against B_step = 60*Np + 64*Nc^3 bytes;
  stages     per-stage mean milliseconds of the same timed steps;
  cpu_baseline   the CPU oracle port (oracle/) on this box's host cores, bounded sample.

--impl reference times the reference's CPU implementation of the same step: the reference is
pure Python + numba + pyFFTW and cannot travel to the GPU box (no /root/reference there, pyFFTW
not installable), so the arm runs the oracle *port* (oracle/pm_oracle.c + scipy.fft, which
tests/test_oracle_golden.py pins bit-for-bit to the reference's own code) with all host threads.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
import types

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "particle-steps/sec"
UNIT = "particle-steps/s"


def measured_hbm_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
VEL_SIGMA = 0.05   # code units; ~0.05 cells of drift per axis per step at a = 0.01 (da*f/(a+da)^2 ~ 1)


def make_particles_torch(n_parts, n_cells, device, seed=38, vel_sigma=VEL_SIGMA):
    """Lattice (zeldovich.py:79-83: row-0 coordinate slowest, +0.5) + uniform(-2,2) jitter
    (zeldovich.py:89-91) wrapped to [0, Nc); Gaussian velocities of rms `vel_sigma` per axis, so
    that ~10 % of the particles change cell in a step as in a running simulation (the reference's
    probed maximum drift is 0.125 cells per step).  Generated on the host with torch's
    CPU generator (deterministic), float32."""
    import torch
    g = torch.Generator().manual_seed(seed)
    res = n_cells / n_parts
    ax = torch.arange(n_parts, dtype=torch.float64) * res + 0.5
    npart = n_parts ** 3
    pos = torch.empty((3, npart), dtype=torch.float32)
    idx = torch.arange(npart, dtype=torch.int64)
    comps = (idx // (n_parts * n_parts), (idx // n_parts) % n_parts, idx % n_parts)
    for d in range(3):
        jit = (torch.rand(npart, generator=g, dtype=torch.float64) * 4.0 - 2.0)
        pos[d] = torch.remainder(ax[comps[d]] + jit, float(n_cells)).to(torch.float32)
    pos.clamp_(max=float(n_cells))
    vel = (torch.randn((3, npart), generator=g, dtype=torch.float32) * vel_sigma)
    return pos, vel


def make_particles_clustered(n_parts, n_cells, seed=38, vel_sigma=VEL_SIGMA):
    """BASELINE configs[4] / SURVEY 8d(iii): the adversarial load for the deposit, the sort and the
    gather -- half of the particles in a Gaussian blob of one cell rms around the box centre (the
    probe behind SURVEY's "13 239 particles in one cell"), half uniform.  Host generator, float32."""
    import torch
    g = torch.Generator().manual_seed(seed)
    npart = n_parts ** 3
    nblob = npart // 2
    pos = torch.empty((3, npart), dtype=torch.float32)
    for d in range(3):
        blob = torch.randn(nblob, generator=g, dtype=torch.float64) + n_cells / 2.0
        uni = torch.rand(npart - nblob, generator=g, dtype=torch.float64) * n_cells
        pos[d] = torch.remainder(torch.cat([blob, uni]), float(n_cells)).to(torch.float32)
    pos.clamp_(max=float(n_cells))
    perm = torch.randperm(npart, generator=g)          # no favourable storage order
    pos = pos[:, perm].contiguous()
    vel = (torch.randn((3, npart), generator=g, dtype=torch.float32) * vel_sigma)
    return pos, vel


def make_particles_slab_gpu(n_parts, n_cells, rank, nranks, dev, seed=38, vel_sigma=VEL_SIGMA):
    """The same IC-like particle load (lattice + uniform(-2,2) jitter), generated on the GPU for ONE
    slab: only lattice planes that can land in the slab are visited, the jitter is a counter-based
    hash of the global particle index so every rank draws the same numbers.  Used for the large
    configurations, where building 10^9 particles on the host of every process is not an option."""
    import torch
    res = n_cells / n_parts
    nzl = n_cells // nranks
    z0 = rank * nzl
    device = f"cuda:{dev}"

    def i64(v):                 # wrap a Python int into the int64 range (two's complement)
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v

    def uniform(idx, stream, salt=0):   # idx: int64 tensor -> float64 in [0, 1)
        x = idx * 3 + stream + i64((seed + 1000003 * salt) * 0x9E3779B97F4A7C15)
        x = (x ^ (x >> 30)) * (-4658895280553007687)          # 0xBF58476D1CE4E5B9 as int64
        x = (x ^ (x >> 27)) * (-7723592293110705685)          # 0x94D049BB133111EB as int64
        x = x ^ (x >> 31)
        return ((x >> 11) & ((1 << 53) - 1)).to(torch.float64) / float(1 << 53)

    iz_lo = int((z0 - 3.0) // res) - 1
    iz_hi = int((z0 + nzl + 3.0) // res) + 2
    izs = torch.arange(iz_lo, iz_hi, device=device, dtype=torch.int64) % n_parts
    if iz_hi - iz_lo >= n_parts:
        izs = torch.arange(n_parts, device=device, dtype=torch.int64)
    iy = torch.arange(n_parts, device=device, dtype=torch.int64)
    out_p, out_v, out_i = [], [], []
    chunk = max(1, min(n_parts, (1 << 25) // max(1, n_parts * izs.numel())))
    for ix0 in range(0, n_parts, chunk):
        ix = torch.arange(ix0, min(n_parts, ix0 + chunk), device=device, dtype=torch.int64)
        gid = ((ix[:, None, None] * n_parts + iy[None, :, None]) * n_parts + izs[None, None, :]).reshape(-1)
        comp = (gid // (n_parts * n_parts), (gid // n_parts) % n_parts, gid % n_parts)
        z = torch.remainder(comp[2].to(torch.float64) * res + 0.5 + uniform(gid, 2) * 4.0 - 2.0,
                            float(n_cells)).to(torch.float32)
        own = (torch.remainder(torch.floor(z).to(torch.int64), n_cells) // nzl) == rank
        gid = gid[own]
        p = torch.empty((3, gid.numel()), dtype=torch.float32, device=device)
        p[2] = z[own]
        for d in (0, 1):
            c = ((gid // (n_parts * n_parts)) if d == 0 else ((gid // n_parts) % n_parts)).to(torch.float64)
            p[d] = torch.remainder(c * res + 0.5 + uniform(gid, d) * 4.0 - 2.0, float(n_cells)).to(torch.float32)
        v = torch.empty_like(p)
        for d in range(3):      # Box-Muller on two more hash streams per axis
            u1 = uniform(gid, d, 1).clamp_(min=1e-300)
            u2 = uniform(gid, d, 2)
            v[d] = (torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2) * vel_sigma).to(torch.float32)
        out_p.append(p)
        out_v.append(v)
        out_i.append(gid.to(torch.int32))
    pos = torch.cat(out_p, dim=1).clamp_(max=float(n_cells)).contiguous()
    vel = torch.cat(out_v, dim=1).contiguous()
    ids = torch.cat(out_i).contiguous()
    return pos, vel, ids


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which
    omp_get_max_threads() (oracle.max_threads) obeys -- the CPU arm must not: it is the reference timed
    on ALL of the box's host cores, whatever launched it (the oracle takes the count explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def workload_label(n_parts, n_cells, particles="ic"):
    """config.workload, derived from the sizes actually run (never a literal)."""
    base = f"{n_parts}^3 particles on {n_cells}^3 mesh, full PM step (CIC deposit + FFT Poisson + gather/kick/drift)"
    which = {(64, 128): "configs[0]", (256, 512): "configs[1]", (512, 1024): "configs[2]",
             (1024, 2048): "configs[3]"}.get((n_parts, n_cells))
    if particles == "clustered":
        return base + ", BASELINE configs[4] clustered microbench load (synthetic blob)"
    if particles == "evolved":
        return base + ", BASELINE configs[4] clustered z=0 snapshot (own ICs evolved by the step itself)"
    return base + (f", BASELINE {which}" if which else "")


def cfg_namespace(n_parts, n_cells, steps_cfg=1000):
    """configure_me defaults (src/configure_me.py:7-40) with the run's sizes."""
    from cosmological_particle_mesh_simulation_b200 import configure_me as cm
    d = {name: entry[0] for name, entry in cm.PARAMETERS.items()}
    d.update(N_PARTS=n_parts, N_CELLS=n_cells, N_CPU=1, STEPS=steps_cfg)
    return types.SimpleNamespace(**d)


def make_particles_evolved(pm, cfg, dev, evolve_steps=None):
    """BASELINE configs[4] / SURVEY 8d(ii): the z = 0 snapshot.  The package's own initial conditions
    (gaussian_random_field -> zeldovich, RANDOM_SEED 38, on the GPU) evolved by the resident step itself
    over the reference's whole schedule (src/pmesh.py:56-63: 999 iterations for STEPS = 1000).
    Returns (pos, vel, a_last, da, n_evolved) with the particles in original order on the device."""
    import torch
    with torch.cuda.device(dev):
        rho0 = pm.gaussian_random_field(device=dev)
        pos, vel = pm.zeldovich(rho0)
        del rho0
    sched = pm.loop_scale_factors(cfg)
    if evolve_steps is not None:
        sched = sched[:int(evolve_steps)]
    mass = (cfg.N_CELLS / cfg.N_PARTS) ** 3
    state = pm.ResidentParticles(pos, vel)
    for a, da in sched:
        state.step(a, da, mass=mass)
    state.store(pos, vel)
    torch.cuda.synchronize()
    a_last, da = sched[-1]
    return pos, vel, a_last, da, len(sched)


def published_ratio(value, n_parts, n_cells):
    """value / the one number the reference publishes for this metric (README.md:16,
    ~1 hour for 999 steps of 256^3 on 512^3 => ~4.7e6 particle-steps/s); null for any other configuration."""
    return value / 4.7e6 if (n_parts, n_cells) == (256, 512) else None


def b_step_bytes(npart, n_cells):
    return 60 * npart + 64 * n_cells ** 3


# algorithmic bytes per launch of each hand-written stage (DESIGN.md "Kernels")
def stage_alg_bytes(npart, n_cells):
    m = n_cells ** 3
    return {
        "keys": 4 * npart,                           # resident path: keys come out of the gather kernel
        "rows": 4 * npart + 4 * n_cells ** 2,        # read sorted keys; write row offsets
        "deposit": 12 * npart + 4 * m,               # SURVEY 8d deposit row
        # The hand-written Poisson solve is FIVE passes over a 4*M-byte array: rows + y
        # forward (stage "fft_r2c": 2 x 8*M), the fused z pass = z forward + Green + z inverse (stage
        # "green": 8*M), y + rows inverse (stage "fft_c2r": 2 x 8*M).  The contract figure B_step still
        # charges the solve 56*M; these are the bytes each STAGE really has to move.
        "green": 8 * m,
        "gather_kick_drift": 48 * npart + 4 * m,     # SURVEY 8d gather row
        "sort": 8 * npart + 8 * npart,               # read keys + previous keys; write sorted keys + order
        "fft_r2c": 16 * m, "fft_c2r": 16 * m,
    }


OWN_STAGES = ("keys", "rows", "deposit", "green", "gather_kick_drift")


def cpu_baseline_sample(n_parts, n_cells, pos_h, vel_h, threads, nsteps=1):
    """The oracle port on this box's host cores; returns (particle-steps/s, seconds)."""
    from oracle import oracle as O
    cfg = O.Config(N_CELLS=n_cells, N_PARTS=n_parts, N_CPU=threads)
    fg = O.fourier_grid(cfg)
    a, da = 0.01, 0.99 / 1000
    mass = (n_cells / n_parts) ** 3
    t0 = time.perf_counter()
    for _ in range(nsteps):
        O.step(pos_h, vel_h, fg, a, da, cfg, mass=mass)
        a += da
    dt = time.perf_counter() - t0
    return pos_h.shape[1] * nsteps / dt, dt


# ------------------------------------------------------------------------------------------------
# arms
# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: oracle port, all host threads, same config/metric.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as O
    threads = host_threads()
    n_parts, n_cells = args.n_parts, args.n_cells
    pos, vel = make_particles_torch(n_parts, n_cells, "cpu")
    pos_h, vel_h = pos.numpy(), vel.numpy()
    npart = pos_h.shape[1]
    # bounded: each step is the full workload; the number of steps is capped by a time budget
    _, t_first = cpu_baseline_sample(n_parts, n_cells, pos_h, vel_h, threads, 1)   # warm-up step
    budget = args.reference_budget_s
    k = max(1, min(args.steps, int(budget / max(t_first, 1e-3))))
    rate, dt = cpu_baseline_sample(n_parts, n_cells, pos_h, vel_h, threads, k)
    sample = f"{k} full steps of {n_parts}^3 on {n_cells}^3 after 1 warm-up step (requested {args.steps}; capped by a {budget:.0f} s budget)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": k, "warmup": 1, "requested_steps": args.steps, "ms_per_step": 1e3 * dt / k,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": published_ratio(rate, n_parts, n_cells),
        "dtype": "f64 (complex128 FFT, float32 state)", "data": "synthetic",
        "config": {"workload": workload_label(n_parts, n_cells),
                   "n_parts": n_parts, "n_cells": n_cells},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference's numba+pyFFTW step (C + scipy.fft complex128); the "
                "reference itself cannot run on this box (no /root/reference, no pyFFTW)",
    }
    print(json.dumps(line), flush=True)


def weak_scaling_record(args, rank, world, dev, dist, pm, comm, transport, chunks, weak_sizes=(512, 1024, 1024, 2048)):
    """BASELINE configs[3] (1024^3 particles on a 2048^3 mesh over all `world` = 8 GPUs) against configs[2]
    (512^3 on 1024^3) on ONE GPU: the same load per GPU, so parallel efficiency = t(1 GPU) / t(8 GPUs)
    (north_star asks for >= 70 %).  A bounded sub-record of the --gpus 8 line: 3 warm-up + 3
    timed steps each.  Every rank enters; rank 0 alone runs the one-GPU part."""
    import torch
    slab = pm.slab
    K, W = 3, 3
    p1, c1, pn, cn = weak_sizes
    rec = {"workload_1gpu": workload_label(p1, c1), "workload_all_gpus": workload_label(pn, cn), "steps": K, "warmup": W}
    ok = torch.ones(1, dtype=torch.int32, device=f"cuda:{dev}")
    err = ""

    def agree():
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(int(ok.item()))

    # ---- one GPU: 512^3 on 1024^3, resident path, on rank 0 ----
    ms1 = 0.0
    try:
        if rank == 0:
            cfg1 = cfg_namespace(p1, c1)
            pm.set_config(cfg1)
            pl, vl, il = make_particles_slab_gpu(p1, c1, 0, 1, dev)
            del il
            st = pm.ResidentParticles(pl, vl)
            del pl, vl
            torch.cuda.empty_cache()
            sched = pm.loop_scale_factors(cfg1)
            side = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(side):
                for i in range(W):
                    st.step(*sched[i], mass=8.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(W, W + K):
                    st.step(*sched[i], mass=8.0)
                e1.record()
                side.synchronize()
                ms1 = e0.elapsed_time(e1) / K
            st.close()
            torch.cuda.empty_cache()
    except Exception as e:      # noqa: BLE001 -- reported in the record, the main line must still be printed
        ok.zero_()
        err = "1-GPU part: " + repr(e)[:200]
    if not agree():
        return dict(rec, error=err or "failed on another rank")
    t = torch.tensor([ms1], dtype=torch.float64, device=f"cuda:{dev}")
    dist.broadcast(t, 0)
    ms1 = float(t.item())
    # ---- all GPUs: 1024^3 on 2048^3, slab path ----
    n_parts, n_cells = pn, cn
    cfg = cfg_namespace(n_parts, n_cells)
    pm.set_config(cfg)
    ranks = []
    try:
        pl, vl, il = make_particles_slab_gpu(n_parts, n_cells, rank, world, dev)
        ranks = [slab.make_rank_from_local(n_cells, pl, vl, il, rank, world, device=dev, total_particles=n_parts ** 3)]
        del pl, vl, il
        torch.cuda.empty_cache()
    except Exception as e:      # noqa: BLE001
        ok.zero_()
        err = "slab set-up: " + repr(e)[:200]
    if not agree():
        for r in ranks:
            r.close()
        return dict(rec, error=err or "failed on another rank")
    tr = "nccl"
    if transport != "nccl" and slab.setup_peers(ranks, comm):
        tr = transport
    aux = tr != "nccl" and slab.setup_ghost_peers(ranks, comm)
    sched = pm.loop_scale_factors(cfg)
    timer = slab.PhaseTimer()
    kw = dict(mass=8.0, cfg=cfg, chunks=chunks or None, transport=tr, ghosts="peer" if aux else "nccl",
              migrate="peer" if aux else "nccl")
    for i in range(W):
        slab.slab_step(ranks, comm, *sched[i], **kw)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + K):
        slab.slab_step(ranks, comm, *sched[i], timer=timer, **kw)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=f"cuda:{dev}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms8 = float(t.item())
    phases = timer.mean_ms()
    tmo = torch.tensor([ranks[0].peer_timeouts() if tr != "nccl" else 0], dtype=torch.int64, device=f"cuda:{dev}")
    dist.all_reduce(tmo, op=dist.ReduceOp.MAX)
    slab.release_peers(ranks, comm)
    for r in ranks:
        r.close()
    torch.cuda.empty_cache()
    peak, _ = measured_hbm_peak()
    bstep = b_step_bytes(n_parts ** 3, n_cells)
    nvlink_bytes = 2 * (4 * n_cells ** 3 / world) * (world - 1) / world
    t_roof = (bstep / world) / (peak * 1e9) + nvlink_bytes / 900e9
    rec.update({"ms_per_step_1gpu": ms1, "ms_per_step_all_gpus": ms8, "n_gpus": world,
                "parallel_efficiency": ms1 / ms8 if ms8 > 0 else None, "target": 0.70,
                "value_all_gpus": n_parts ** 3 / (ms8 * 1e-3), "unit": UNIT,
                "fft_transport": tr, "ghost_planes": kw["ghosts"], "migration": kw["migrate"],
                "phases_ms_rank0": phases, "peer_flag_timeouts": int(tmo.item()),
                "roofline_step": {"t_roof_ms": 1e3 * t_roof, "frac": 1e3 * t_roof / ms8,
                                  "formula": "(60*Np+64*Nc^3)/P/hbm + 2*(4*Nc^3/P)*(P-1)/P/900e9 (SURVEY 8e, serial bound)"}})
    return rec


def run_slab(args, rank, world, local_rank):
    """N > 1: the 256^3/512^3 step slab-decomposed over the GPUs of the box (strong scaling)."""
    import torch
    import torch.distributed as dist
    import cosmological_particle_mesh_simulation_b200 as pm
    slab = pm.slab
    torch.cuda.set_device(local_rank)
    dev = local_rank
    # NCCL prints "NCCL version ..." on stdout at NCCL_DEBUG=VERSION; stdout must hold only the JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    n_parts, n_cells = args.n_parts, args.n_cells
    cfg = cfg_namespace(n_parts, n_cells)
    pm.set_config(cfg)
    npart = n_parts ** 3
    mass = (n_cells / n_parts) ** 3
    comm = slab.DistComm()
    def build_ranks():
        if args.particles == "zeldovich":
            # the package's own initial conditions, generated slab by slab (slab_ic.py): no rank holds the lattice
            out = pm.slab_ic.make_ranks_from_ic(comm, cfg=cfg, device=dev)
            cnt = torch.tensor([out[0].count], dtype=torch.int64, device=f"cuda:{dev}")
            dist.all_reduce(cnt)
            assert int(cnt.item()) == npart, (int(cnt.item()), npart)
            desc = "gaussian_random_field + zeldovich generated per slab on the GPUs (Philox, RANDOM_SEED 38)"
        elif n_parts > 256:
            # large configurations: every rank generates its own slab on its GPU
            pl, vl, il = make_particles_slab_gpu(n_parts, n_cells, rank, world, dev)
            out = [slab.make_rank_from_local(n_cells, pl, vl, il, rank, world, device=dev, total_particles=npart)]
            cnt = torch.tensor([pl.shape[1]], dtype=torch.int64, device=f"cuda:{dev}")
            dist.all_reduce(cnt)
            assert int(cnt.item()) == npart, (int(cnt.item()), npart)
            desc = ("lattice + uniform(-2,2) jitter, Gaussian velocities rms %g (counter-based hash, seed 38), "
                    "generated per slab on the GPU" % VEL_SIGMA)
        else:
            pos_h, vel_h = make_particles_torch(n_parts, n_cells, "cpu")
            pos, vel = pos_h.to(f"cuda:{dev}"), vel_h.to(f"cuda:{dev}")
            out = slab.make_ranks(n_cells, pos, vel, comm, device=dev)
            desc = "lattice + uniform(-2,2) jitter, Gaussian velocities rms %g, seed 38" % VEL_SIGMA
        torch.cuda.empty_cache()
        return out, desc

    ranks, particles_desc = build_ranks()
    sched = pm.loop_scale_factors(cfg)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    step_i = 0

    def warm_up(transport):
        nonlocal step_i
        for _ in range(W):
            a, da = sched[step_i % len(sched)]
            slab.slab_step(ranks, comm, a, da, mass=mass, cfg=cfg, chunks=args.chunks or None, transport=transport,
                           ghosts=ghosts, migrate=migrate)
            step_i += 1
        barrier()

    # FFT transposes through peer memory (CUDA IPC + NVLink stores/loads) unless --transport nccl.
    # setup_peers() ends with a flag handshake through the mapped memory and all ranks agree on the
    # outcome; if the set-up fails, or a flag wait times out during warm-up on any rank, every rank
    # rebuilds its state and runs the NCCL all-to-all path instead -- and the JSON line says which.
    transport, transport_note, ghosts, migrate = "nccl", "", "nccl", "nccl"
    if args.transport != "nccl":
        transport = (args.transport if args.transport in ("peer", "fused2") else "fused") if slab.setup_peers(ranks, comm) else "nccl"
        if transport == "nccl":
            transport_note = "peer-memory set-up failed; "
    if transport != "nccl" and (args.ghosts != "nccl" or args.migrate != "nccl"):
        # ghost planes and particle migration through the same peer mappings (default when they can be set up)
        if slab.setup_ghost_peers(ranks, comm):
            ghosts = "nccl" if args.ghosts == "nccl" else "peer"
            migrate = "nccl" if args.migrate == "nccl" else "peer"
        elif "peer" in (args.ghosts, args.migrate):
            raise RuntimeError("--ghosts/--migrate peer need a working peer-memory set-up")
    warm_up(transport)
    if transport != "nccl":
        bad = torch.tensor([ranks[0].peer_timeouts()], dtype=torch.int64, device=f"cuda:{dev}")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            transport, transport_note, ghosts, migrate = "nccl", "peer-memory flag wait timed out in warm-up; ", "nccl", "nccl"
            slab.release_peers(ranks, comm)
            for r in ranks:
                r.close()
            ranks, particles_desc = build_ranks()
            step_i = 0
            warm_up(transport)
    if args.transport in ("peer", "fused", "fused2") and transport != args.transport:
        raise RuntimeError(f"--transport {args.transport}: " + transport_note)
    # sanity on the distributed state: total mass of the last deposit == Np * mass
    msum = ranks[0].buf["RHO"].sum(dtype=torch.float64).reshape(1)
    dist.all_reduce(msum)
    mass_err = abs(float(msum.item()) / (npart * mass) - 1.0)
    assert mass_err < 1e-5, mass_err
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.3)
    timer = slab.PhaseTimer()
    launches0 = pm.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        a, da = sched[step_i % len(sched)]
        slab.slab_step(ranks, comm, a, da, mass=mass, cfg=cfg, timer=timer, chunks=args.chunks or None,
                       transport=transport, ghosts=ghosts, migrate=migrate)
        step_i += 1
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = pm.launch_count() - launches0
    clocks = sampler.stop()
    phases = timer.mean_ms()
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{dev}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = npart * K / (ms_total * 1e-3)
    # a flag wait that gave up (2 s) inside the timed steps would have let a kernel read stale data:
    # report it instead of a silently wrong number (0 on a healthy run; the max over ranks)
    tmo = torch.tensor([ranks[0].peer_timeouts() if transport != "nccl" else 0], dtype=torch.int64, device=f"cuda:{dev}")
    dist.all_reduce(tmo, op=dist.ReduceOp.MAX)
    peer_timeouts = int(tmo.item())

    e2e_value, ke = None, 0
    t = torch.zeros(3, dtype=torch.float64, device=f"cuda:{dev}")
    if not args.no_e2e:
        # e2e: the rank's particles live in pinned host memory; every step uploads them, runs the
        # slab step and downloads the rank's (migrated) particles again
        sr = ranks[0]
        p, v, ids = sr.export()
        cap = sr.np_capacity
        # pinned host state of this rank, allocated once (capacity of the plan): the timed steps only copy
        hp_flat = torch.empty(3 * cap, dtype=torch.float32).pin_memory()
        hv_flat = torch.empty(3 * cap, dtype=torch.float32).pin_memory()
        hi = torch.empty(cap, dtype=torch.int32).pin_memory()
        n_host = p.shape[1]

        def host_views(n):      # contiguous [3, n] views, so every copy is one plain DMA
            return hp_flat[:3 * n].view(3, n), hv_flat[:3 * n].view(3, n), hi[:n]

        hp, hv, hid = host_views(n_host)
        hp.copy_(p); hv.copy_(v); hid.copy_(ids)
        torch.cuda.synchronize()
        ke = max(3, min(K, 10))
        h2d = d2h = 0

        def host_step():
            nonlocal n_host, h2d, d2h, step_i
            a, da = sched[step_i % len(sched)]
            step_i += 1
            hp, hv, hid = host_views(n_host)
            dp = hp.to(f"cuda:{dev}", non_blocking=True)
            dv = hv.to(f"cuda:{dev}", non_blocking=True)
            di = hid.to(f"cuda:{dev}", non_blocking=True)
            h2d += n_host * 4 * 7
            sr.load(dp, dv, di)
            slab.slab_step(ranks, comm, a, da, mass=mass, cfg=cfg, transport=transport, ghosts=ghosts, migrate=migrate)
            p, v, ids = sr.export()
            n_host = p.shape[1]
            hp, hv, hid = host_views(n_host)
            hp.copy_(p, non_blocking=True)
            hv.copy_(v, non_blocking=True)
            hid.copy_(ids, non_blocking=True)
            torch.cuda.synchronize()          # the step's results are in host memory
            d2h += n_host * 4 * 7

        for _ in range(2):
            host_step()
        barrier()
        h2d = d2h = 0
        t0 = time.perf_counter()
        for _ in range(ke):
            host_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s, float(h2d), float(d2h)], dtype=torch.float64, device=f"cuda:{dev}")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e_value = npart * ke / float(tmax[0].item())

    weak = None
    weak_sizes = tuple(int(v) for v in args.weak_scaling_sizes.split(",")) if args.weak_scaling_sizes else None
    if ((world == 8 and (n_parts, n_cells) == (256, 512)) or weak_sizes) and not args.no_weak_scaling:
        # free the strong-scaling state first: configs[3] needs the memory
        slab.release_peers(ranks, comm)
        for r in ranks:
            r.close()
        ranks = []
        torch.cuda.empty_cache()
        weak = weak_scaling_record(args, rank, world, dev, dist, pm, comm, transport, args.chunks,
                                   weak_sizes or (512, 1024, 1024, 2048))
        pm.set_config(cfg)
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        bstep = b_step_bytes(npart, n_cells)
        nvlink_bytes = 2 * (4 * n_cells ** 3 / world) * (world - 1) / world   # per GPU per step
        t_roof = (bstep / world) / (peak * 1e9) + nvlink_bytes / 900e9
        dominant = max(phases, key=phases.get)
        chunks = args.chunks or slab.default_chunks(n_cells, world, transport)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": published_ratio(value, n_parts, n_cells), "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(n_parts, n_cells),
                       "n_parts": n_parts, "n_cells": n_cells,
                       "particles": particles_desc,
                       "l2": "inputs larger than L2",
                       "parallelism": f"slab decomposition along z over {world} GPUs: ghost planes by "
                                      + ("peer-memory stores" if ghosts == "peer" else "NCCL send/recv") + ", "
                                      + {"fused": "FFT transposes fused into the y passes (stores into / loads from the peers' "
                                                  "z-pass arrays over NVLink, CUDA IPC, flag-word barriers), ",
                                         "fused2": "FFT transposes fused into the y passes, y passes and z passes on two streams (experimental), ",
                                         "peer": "FFT transposes by peer-memory copy kernels over NVLink (CUDA IPC, flag-word barriers), ",
                                         "nccl": "FFT transposes by NCCL all-to-all, "}[transport]
                                      + f"pipelined in {chunks} kx chunks" + ("" if transport == "fused" else " on two streams")
                                      + (", particle migration through peer memory (count matrix + records, one host read)"
                                         if migrate == "peer" else ", all-to-all-v particle migration"),
                       "fft_transport": transport_note + transport, "ghost_planes": ghosts, "migration": migrate},
            "clocks": clocks,
            "e2e": None if e2e_value is None else {
                    "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": float(t[1].item()) / ke,
                    "d2h_bytes_per_step": float(t[2].item()) / ke, "steps": ke,
                    "api": "per rank: pinned host pos+vel+ids -> pm_slab_load -> slab step -> pm_slab_export -> host"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm+nvlink", "kernel": "phase:" + dominant, "achieved": None, "peak": peak,
                         "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                         "ms_per_launch": phases[dominant]},
            "roofline_step": {"bound": "hbm+nvlink", "t_roof_ms": 1e3 * t_roof, "frac": 1e3 * t_roof / ms_per_step,
                              "formula": "(60*Np+64*Nc^3)/P/hbm + 2*(4*Nc^3/P)*(P-1)/P/900e9 (SURVEY 8e, serial bound)"},
            "phases_ms_rank0": phases,
            "mass_conservation_rel_err": mass_err,
            "peer_flag_timeouts": peer_timeouts,
            "weak_scaling": weak,
        }
        print(json.dumps(line), flush=True)
    slab.release_peers(ranks, comm)
    for r in ranks:
        r.close()
    dist.destroy_process_group()


def run_ours(args, rank, world, local_rank):
    if world > 1:
        return run_slab(args, rank, world, local_rank)
    import torch
    import cosmological_particle_mesh_simulation_b200 as pm
    rt = pm._runtime

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product arm)")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))

    n_parts, n_cells = args.n_parts, args.n_cells
    cfg = cfg_namespace(n_parts, n_cells)
    pm.set_config(cfg)
    npart = n_parts ** 3
    mass = (n_cells / n_parts) ** 3
    if args.particles == "clustered":
        pos_h, vel_h = make_particles_clustered(n_parts, n_cells)
        particles_desc = ("clustered microbench load (BASELINE configs[4]): half in a 1-cell-rms Gaussian blob at the "
                          "box centre, half uniform, shuffled; Gaussian velocities rms %g, seed 38" % VEL_SIGMA)
    elif args.particles == "evolved":
        pos, vel, a_last, da_last, n_ev = make_particles_evolved(pm, cfg, dev, args.evolve_steps or None)
        pos_h, vel_h = pos.cpu(), vel.cpu()
        particles_desc = ("z=0-like snapshot (BASELINE configs[4]): own Zel'dovich ICs (RANDOM_SEED 38) evolved by "
                          "%d resident steps of the reference schedule to a = %.4f" % (n_ev, a_last + da_last))
    elif args.particles == "zeldovich":
        with torch.cuda.device(dev):
            pos, vel = pm.zeldovich(pm.gaussian_random_field(device=dev))
        pos_h, vel_h = pos.cpu(), vel.cpu()
        particles_desc = "the package's own initial conditions (gaussian_random_field + zeldovich on the GPU, RANDOM_SEED 38)"
    else:
        pos_h, vel_h = make_particles_torch(n_parts, n_cells, "cpu")
        particles_desc = "lattice + uniform(-2,2) jitter, Gaussian velocities rms %g, seed 38" % VEL_SIGMA
    pos, vel = pos_h.to(f"cuda:{dev}"), vel_h.to(f"cuda:{dev}")
    sched = pm.loop_scale_factors(cfg)
    if args.particles == "evolved":
        sched = [(a_last, da_last)]          # keep stepping at the late-time scale factor

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    step_i = 0
    state = pm.ResidentParticles(pos, vel)   # state resident in HBM, cell-ordered between steps
    plan = state.plan
    for _ in range(W):
        a, da = sched[step_i % len(sched)]
        state.step(a, da, mass=mass)
        step_i += 1
    barrier()

    # ---- timed region: K resident steps, CUDA events on the launching stream ----
    # (a side stream: the legacy default stream cannot be captured, and the steady-state step replays as a
    # CUDA graph -- include/pmstep.h, pm_plan_set_graph)
    import ctypes
    import numpy as np
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.3)
    with torch.cuda.stream(side):
        for _ in range(2):                       # the first steps on this stream capture the two graphs
            a, da = sched[step_i % len(sched)]
            state.step(a, da, mass=mass)
            step_i += 1
        barrier()
        replays0 = int(rt.lib().pm_plan_graph_replays(plan.handle))
        launches0 = pm.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(K):
            a, da = sched[step_i % len(sched)]
            state.step(a, da, mass=mass)
            step_i += 1
        ev1.record()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        launches_eager = pm.launch_count() - launches0
        graph_replays = int(rt.lib().pm_plan_graph_replays(plan.handle)) - replays0
        # ---- second pass, eager, with CUDA events between the stages: the per-stage breakdown ----
        rt.check(rt.lib().pm_plan_profile_begin(plan.handle, K), "profile_begin")
        launches0 = pm.launch_count()
        evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evp0.record()
        for _ in range(K):
            a, da = sched[step_i % len(sched)]
            state.step(a, da, mass=mass)
            step_i += 1
        evp1.record()
        barrier()
        ms_profiled = evp0.elapsed_time(evp1) / K
        launches = pm.launch_count() - launches0          # kernels of one eager pass = kernels inside each replay
    torch.cuda.current_stream(dev).wait_stream(side)
    sort_n, sort_movers, sort_mode = state.sort_stats()
    block_stats = state.block_stats()
    gi = state.gather_items()
    gather_items = {"heavy": gi[0], "light": gi[1], "overflow": gi[2],
                    "note": "work list of the tiled gather: columns above twice the mean load are cut into pieces and dispatched first"}
    fft_sync_errors = int(rt.lib().pm_plan_fft_sync_errors(plan.handle))
    nst = len(rt.STAGE_NAMES)
    buf = np.zeros((K, nst), dtype=np.float32)
    nrec = ctypes.c_int(0)
    rt.check(rt.lib().pm_plan_profile_read(plan.handle, buf.ctypes.data, ctypes.byref(nrec)), "profile_read")
    rt.check(rt.lib().pm_plan_profile_begin(plan.handle, 0), "profile_end")
    clocks = sampler.stop()
    stage_ms = {n: float(buf[:nrec.value, i].mean()) for i, n in enumerate(rt.STAGE_NAMES)}

    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = npart * world * K / (ms_total * 1e-3)   # replicas until the slab path lands

    # ---- e2e: host-buffer C-ABI call, pinned host memory, copies inside the timed region ----
    e2e_value, ke = None, 0
    if not args.no_e2e:
        state.store(pos, vel)
        ph, vh = pos.cpu().pin_memory(), vel.cpu().pin_memory()
        ke = max(3, min(K, 10))
        for _ in range(2):
            a, da = sched[step_i % len(sched)]
            pm.step_host(ph, vh, a, da, mass=mass, device=dev)
            step_i += 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            a, da = sched[step_i % len(sched)]
            pm.step_host(ph, vh, a, da, mass=mass, device=dev)
            step_i += 1
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{dev}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_value = npart * world * ke / e2e_s

    # ---- drop-in leg: the reference's own loop body (src/pmesh.py:60-61), statement for statement ----
    dropin = None
    if not args.no_e2e:
        dropin = {}
        state.store(pos, vel)
        fgrid = pm.fourier_grid()
        pd, vd = pos.clone(), vel.clone()

        def loop_body(p_, v_, a_, da_):
            rho_ = pm.density(p_, mass)                                          # pmesh.py:60
            return pm.advance_time(rho_, p_, v_, fgrid, a_, da_)                  # pmesh.py:61

        for _ in range(3):                      # first call: stateless + session set-up; then the resident session
            a, da = sched[step_i % len(sched)]
            loop_body(pd, vd, a, da)
            step_i += 1
        barrier()
        kd = max(3, min(K, 20))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(kd):
            a, da = sched[step_i % len(sched)]
            loop_body(pd, vd, a, da)
            step_i += 1
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / kd
        dropin["cuda_tensors"] = {"value": npart / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": kd,
                                  "api": "rho = density(pos, mass); pos, vel = advance_time(rho, pos, vel, fgrid, a, da) on CUDA "
                                         "tensors: resident session behind the reference's signatures (_session.py)"}
        pm.forget_resident()
        # the same loop with the write-back deferred (set_resident_dropin("lazy"), _session.py): names rebound to
        # the returned handles exactly as src/pmesh.py:61 does; one read of the result after the last step
        pm.set_resident_dropin("lazy")
        try:
            pl_, vl_ = pos.clone(), vel.clone()
            for _ in range(3):
                a, da = sched[step_i % len(sched)]
                rho_ = pm.density(pl_, mass)
                pl_, vl_ = pm.advance_time(rho_, pl_, vl_, fgrid, a, da)
                step_i += 1
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(kd):
                a, da = sched[step_i % len(sched)]
                rho_ = pm.density(pl_, mass)                                      # pmesh.py:60
                pl_, vl_ = pm.advance_time(rho_, pl_, vl_, fgrid, a, da)          # pmesh.py:61
                step_i += 1
            pm.sync_particles()                   # the deferred write-back, once, inside the timed region
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1) / kd
            dropin["cuda_tensors_lazy"] = {"value": npart / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": kd,
                                           "api": "the same loop with set_resident_dropin('lazy'): advance_time returns handles over the "
                                                  "caller's storage, the un-permute into it runs once when they are read (here: after "
                                                  "the last step, inside the timed region)"}
            del pl_, vl_, rho_
        finally:
            pm.forget_resident()
            pm.set_resident_dropin(True)
        del pd, vd
        # NumPy in / NumPy out, as the reference's driver holds its state: every call crosses PCIe
        pn, vn = pos.cpu().numpy().copy(), vel.cpu().numpy().copy()
        kn = 3
        a, da = sched[step_i % len(sched)]
        loop_body(pn, vn, a, da)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(kn):
            a, da = sched[step_i % len(sched)]
            loop_body(pn, vn, a, da)
            step_i += 1
        torch.cuda.synchronize()
        msn = 1e3 * (time.perf_counter() - t0) / kn
        dropin["numpy"] = {"value": npart / (msn * 1e-3), "unit": UNIT, "ms_per_step": msn, "steps": kn,
                           "h2d_bytes_per_step": 2 * 12 * npart + 12 * npart + 4 * n_cells ** 3,
                           "d2h_bytes_per_step": 4 * n_cells ** 3 + 24 * npart,
                           "api": "the same two calls on NumPy arrays (pageable host memory; the density mesh is returned to the host too)"}
        del pn, vn

    if rank != 0:
        return
    peak, peak_src = measured_hbm_peak()
    alg = stage_alg_bytes(npart, n_cells)
    dominant = max(OWN_STAGES, key=lambda n: stage_ms[n])
    ach = alg[dominant] / (stage_ms[dominant] * 1e-3) / 1e9
    bstep = b_step_bytes(npart, n_cells)
    ach_step = bstep / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")   # dram bytes per launch from ncu --set full
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dominant)
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if world == 1 else "weak",
        "vs_baseline": published_ratio(value, n_parts, n_cells) if args.particles == "ic" else None,
        "particles_kind": args.particles,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(n_parts, n_cells, args.particles),
                   "n_parts": n_parts, "n_cells": n_cells,
                   "particles": particles_desc,
                   "sort": {"mode": sort_mode, "mover_fraction_last_step": sort_movers / max(sort_n, 1)},
                   "gather_blocks": block_stats,
                   "gather_items": gather_items,
                   "fft": {"fused_plane_passes": os.environ.get("PM_FFT_FUSE", "0") == "1",
                           "kernels": ("radix-8.8.8" if os.environ.get("PM_FFT_V2", "1") == "0" else
                                       "two-stage" if os.environ.get("PM_FFT_ZMIX", "1") == "0" else
                                       "two-stage rows + %s y + radix-8.8.8 fused z" %
                                       ("two-stage" if os.environ.get("PM_FFT_V3", "2") == "0" else "pipelined two-stage")),
                           "sync_errors": fft_sync_errors},
                   "l2": "inputs larger than L2 (201 MB particles rows, 537 MB meshes vs 126 MB L2)",
                   "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas (slab path pending)"},
        "clocks": clocks,
        "e2e": None if e2e_value is None else {
                "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24 * npart,
                "d2h_bytes_per_step": 24 * npart, "steps": ke,
                "api": "pm_step_host (pinned host pos+vel in, pos+vel out; density stays on device, "
                       "as with the reference defaults SAVE_DENSITY=False, PLOT_*=False)"},
        "e2e_dropin": dropin,
        "gpu_launches": int(launches),
        "graph": {"replays_in_timed_region": graph_replays, "eager_launches_in_timed_region": int(launches_eager),
                  "note": "steady-state steps replay as a CUDA graph; gpu_launches counts the kernels of the same K steps "
                          "run eagerly in the profiled second pass (one replay launches the same kernels)"},
        "ms_per_step_profiled_pass": ms_profiled,
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg[dominant], "ms_per_launch": stage_ms[dominant]},
        "roofline_step": {"bound": "hbm", "achieved": ach_step, "peak": peak, "unit": "GB/s",
                          "frac": ach_step / peak, "algorithmic_bytes_per_step": bstep,
                          "formula": "60*Np + 64*Nc^3 (SURVEY 8d)"},
        "stages_ms": stage_ms,
        "stages_note": "CUDA events between the stages of a second, eager pass over the same K steps (the timed pass replays graphs)",
        "stage_frac_of_peak": {n: alg[n] / (stage_ms[n] * 1e-3) / 1e9 / peak for n in stage_ms if stage_ms[n] > 0.01},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        pc, vc = pos_h.numpy().copy(), vel_h.numpy().copy()
        rate, dt = cpu_baseline_sample(n_parts, n_cells, pc, vc, threads, 1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"1 full step of {n_parts}^3 on {n_cells}^3 ({dt:.1f} s), oracle port "
                                          "(C particle loops + scipy.fft complex128)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-parts", type=int, default=256)
    ap.add_argument("--n-cells", type=int, default=512)
    ap.add_argument("--particles", default="ic", choices=["ic", "clustered", "evolved", "zeldovich"],
                    help="N=1 workload: IC-like lattice+jitter (default, the metric's configuration); the synthetic clustered "
                         "microbench load of BASELINE configs[4]; or `evolved`, the z=0 snapshot obtained by running the "
                         "package's own ICs through the whole schedule (per-stage times show deposit/sort/gather under skew)")
    ap.add_argument("--evolve-steps", type=int, default=0, help="--particles evolved: steps to evolve (0 = the whole schedule)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="multi-GPU exploration runs: skip the host-buffer leg")
    ap.add_argument("--no-weak-scaling", action="store_true",
                    help="--gpus 8: skip the weak_scaling sub-record (configs[3] on 8 GPUs against configs[2] on one)")
    ap.add_argument("--weak-scaling-sizes", default="",
                    help="testing: 'p1,c1,pn,cn' runs the weak_scaling sub-record at any --gpus with these sizes "
                         "(one GPU: p1^3 on c1^3; all GPUs: pn^3 on cn^3)")
    ap.add_argument("--chunks", type=int, default=0, help="kx chunks of the distributed FFT pipeline (0 = auto)")
    ap.add_argument("--transport", default="auto", choices=["auto", "fused", "fused2", "peer", "nccl"],
                    help="FFT transposes of the multi-GPU path: peer-memory copy kernels or NCCL all-to-all")
    ap.add_argument("--ghosts", default="auto", choices=["auto", "nccl", "peer"],
                    help="multi-GPU ghost planes: stores into the neighbours' memory (default when the peer set-up works) or NCCL send/recv")
    ap.add_argument("--migrate", default="auto", choices=["auto", "nccl", "peer"],
                    help="multi-GPU particle migration: through peer memory (default when the peer set-up works) or NCCL all-to-all-v")
    ap.add_argument("--reference-budget-s", type=float, default=90.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
