"""Device time of the slab row passes of one rank (local kernels), two-stage vs radix-8 kernels.
    python scratch/rows_time.py N_CELLS NRANKS"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n, P = int(sys.argv[1]), int(sys.argv[2])
for v2 in ("1", "0"):
    os.environ["PM_FFT_V2"] = v2
    import cosmological_particle_mesh_simulation_b200 as pm
    r = pm.slab.SlabRank(n, 4096, 0, 0, P)
    r.buf["RHO"].uniform_()
    for _ in range(2):
        r.fft_rows_forward(); r.fft_rows_inverse()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    K = 5
    tf = ti = 0.0
    for _ in range(K):
        ev[0].record(); r.fft_rows_forward(); ev[1].record(); r.fft_rows_inverse(); ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1]); ti += ev[1].elapsed_time(ev[2])
    gb = 2 * r.buf["RHO"].numel() * 4 / 1e9
    print(f"{n}^3 mesh, {n // P} planes, PM_FFT_V2={v2}: rows forward {tf / K:.3f} ms ({gb / (tf / K) * 1e3:.0f} GB/s), inverse {ti / K:.3f} ms ({gb / (ti / K) * 1e3:.0f} GB/s)", flush=True)
    r.close()
