"""Diagnostic: resident step with a segmented deposit vs the stateless step, per golden fixture."""
import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import cosmological_particle_mesh_simulation_b200 as pm
from oracle import oracle as O
sys.path.insert(0, "tests")
from test_gpu_parity import load_case, cfg_ns, dev, rel_l2, rel_l2_periodic
gd = "tests/golden"
for xseg in (8, 16):
    for name in ["free16", "clustered32", "g32_step2", "g16_free10"]:
        g, cfg = load_case(gd, name)
        pm.set_config(cfg_ns(cfg))
        pm.release_plans()
        os.environ["PM_DEPOSIT_XSEG"] = str(xseg)
        p1, v1 = dev(g["pos0"]), dev(g["vel0"])
        st = pm.ResidentParticles(p1, v1)
        st.step(float(g["a_list"][0]), float(g["da"]), mass=float(g["mass"]))
        st.store(p1, v1)
        rho1 = pm.density(dev(g["pos0"]), float(g["mass"]))
        pm.release_plans()
        del os.environ["PM_DEPOSIT_XSEG"]
        p2, v2 = dev(g["pos0"]), dev(g["vel0"])
        pm.step(p2, v2, float(g["a_list"][0]), float(g["da"]), mass=float(g["mass"]))
        rho2 = pm.density(dev(g["pos0"]), float(g["mass"]))
        ndiff = int((rho1 != rho2).sum())
        print(xseg, name, "pos", rel_l2_periodic(p1.cpu().numpy(), p2.cpu().numpy(), cfg.N_CELLS),
              "vel", rel_l2(v1.cpu().numpy(), v2.cpu().numpy()), "rho", rel_l2(rho1.cpu().numpy(), rho2.cpu().numpy()),
              "cells differing", ndiff, "rho max", float(rho2.max()), flush=True)
