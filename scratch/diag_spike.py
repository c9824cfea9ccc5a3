"""Per-step parity of the free-running loop on every golden fixture, with the float32 transforms (own FFT,
cuFFT) and with the float64 diagnostic backend: which part of the residue is transform precision?
    PM_FFT_BACKEND={own,cufft,f64} python scratch/diag_spike.py"""
import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
import cosmological_particle_mesh_simulation_b200 as pm

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
def relp(a, b, n):
    d = (np.asarray(a, np.float64) - np.asarray(b, np.float64) + n / 2) % n - n / 2
    return np.linalg.norm(d) / np.linalg.norm(np.asarray(b, np.float64))

gd = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name in ["g16_free10", "g32_step2", "g12_nonpow2", "clustered32", "free16", "free32"]:
    g = np.load(os.path.join(gd, name + ".npz"))
    cfg = O.Config(N_CELLS=int(g["n_cells"]), N_PARTS=int(g["n_parts"]), STEPS=int(g["steps_cfg"]))
    pm.release_plans()
    pm.set_config(types.SimpleNamespace(**cfg.__dict__))
    n = cfg.N_CELLS
    pos, vel = torch.from_numpy(g["pos0"]).cuda(), torch.from_numpy(g["vel0"]).cuda()
    fg = pm.fourier_grid(); da = float(g["da"])
    out = []
    for s, a in enumerate(g["a_list"]):
        rho = pm.density(pos, float(g["mass"]))
        er = rel(rho.cpu().numpy(), g[f"rho_{s}"]) if f"rho_{s}" in g else float("nan")
        pm.advance_time(rho, pos, vel, fg, float(a), da)
        if f"pos_{s+1}" in g:
            out.append((s, er, relp(pos.cpu().numpy(), g[f"pos_{s+1}"], n), rel(vel.cpu().numpy(), g[f"vel_{s+1}"])))
    print(os.environ.get("PM_FFT_BACKEND", "own"), name, " ".join("s%d rho %.1e pos %.1e vel %.1e |" % t for t in out), flush=True)
