"""Wall time of the slab-by-slab initial conditions (slab_ic.slab_initial_conditions), P ranks in one
process on one GPU (LocalComm), against the single-GPU generator at the same size.  Usage:
python scratch/time_slab_ic.py N_PARTS P [P ...]"""
import json
import sys
import time
import types

import torch

sys.path.insert(0, ".")
import cosmological_particle_mesh_simulation_b200 as pm  # noqa: E402
from cosmological_particle_mesh_simulation_b200 import configure_me as cm  # noqa: E402

n = int(sys.argv[1])
d = {k: getattr(cm, k) for k in dir(cm) if k.isupper()}
d.update(N_PARTS=n, N_CELLS=2 * n)
cfg = types.SimpleNamespace(**d)
pm.set_config(cfg)
out = {"n_parts": n, "n_cells": 2 * n}
import importlib  # noqa: E402
G = importlib.import_module("cosmological_particle_mesh_simulation_b200.gaussian_random_field")
Z = importlib.import_module("cosmological_particle_mesh_simulation_b200.zeldovich")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pos, vel = Z.zeldovich(G.gaussian_random_field())
    torch.cuda.synchronize(); out["single_gpu_s"] = time.perf_counter() - t0
for P in [int(x) for x in sys.argv[2:]]:
    comm = pm.slab.LocalComm(P)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        parts = pm.slab_ic.slab_initial_conditions(comm)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    moved = 0
    for r, (p, v, i) in enumerate(parts):
        nl = n // P
        moved += int(((i.long() % n) // nl != r).sum())
    same = sum(int((p == pos[:, i.long()]).sum()) for p, v, i in parts)
    out[f"slab_P{P}_s"] = dt
    out[f"slab_P{P}_routed_fraction"] = moved / n ** 3
    out[f"slab_P{P}_identical_positions"] = same / (3 * n ** 3)
    out[f"slab_P{P}_peak_GB"] = torch.cuda.max_memory_allocated() / 1e9
    del parts
print(json.dumps(out))
