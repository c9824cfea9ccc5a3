"""Parity of the CUDA path (through the C ABI of include/pmstep.h, via the Python drop-in modules)
against the CPU oracle and the golden fixtures.  Run on the B200 box:  pytest -m gpu

Bars (BASELINE.json north_star):
  * cell keys, sort order: BIT-EXACT;
  * gather+kick+drift given the same phi: BIT-EXACT positions and velocities (the kernel repeats
    the reference's float32/float64 rounding points);
  * density, potential, accelerations, positions, velocities: relative L2 <= 1e-5 per step over a
    10-step free-running horizon;
  * P(k) within 0.1 % after a full 64^3/128^3 run.
"""
import os
import types

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

REL_L2 = 1e-5  # north_star tolerance


@pytest.fixture(scope="module")
def pm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cosmological_particle_mesh_simulation_b200 as pm
    yield pm
    pm.set_config(None)
    pm.release_plans()


def cfg_ns(cfg: O.Config):
    return types.SimpleNamespace(**cfg.__dict__)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def rel_l2_periodic(a, b, n):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = (a - b + n / 2) % n - n / 2
    return np.linalg.norm(d) / np.linalg.norm(b)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def load_case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = O.Config(N_CELLS=int(g["n_cells"]), N_PARTS=int(g["n_parts"]), STEPS=int(g["steps_cfg"]))
    return g, cfg


CASES = ["g16_free10", "g32_step2", "g12_nonpow2", "clustered32", "free16", "free32"]
# Stress fixtures: a particle with TWO coordinates == N_CELLS deposits ~Nc^2*m into single cells of
# a tiny mesh (squared).  That spike carries >95 % of ||phi||, so anything that perturbs phi
# RELATIVE TO ITS NORM is amplified ~1e3-fold in every other particle's force: the float32 transform
# noise (cuFFT's or ours).  They stay bit-exact where the arithmetic is shared (keys, sort, gather given
# phi), and the free-running tolerance on them with float32 transforms is that measured floor, not the
# north_star's 1e-5 -- which the realistic fixtures free16/free32/clustered32 and the 64^3 run meet.
# That this is transform precision and nothing else is a TEST, not a claim:
# test_spike_fixtures_meet_1e_5_when_only_the_transform_precision_changes runs the same loop with the
# float64 diagnostic transforms (pm_plan_set_fft_backend 2) and requires 1e-5 at every recorded step.
SPIKE_CASES = {"g16_free10": 1e-3, "g32_step2": 1e-4, "g12_nonpow2": 1e-3}


# ----------------------------------------------------------------------------- keys and sort
@pytest.mark.parametrize("name", CASES)
def test_cell_keys_and_sort_order_bit_exact(pm, golden_dir, name):
    import ctypes
    g, cfg = load_case(golden_dir, name)
    pm.set_config(cfg_ns(cfg))
    rt = pm._runtime
    for key in ["pos0", "pos_1", "pos_2"]:
        if key not in g:
            continue
        pos = g[key]
        npart = pos.shape[1]
        plan = rt.get_plan(cfg.N_CELLS, npart, 0)
        p = dev(pos)
        keys = torch.empty(npart, dtype=torch.int32, device="cuda")
        ks = torch.empty_like(keys)
        order = torch.empty_like(keys)
        rt.check(rt.lib().pm_cell_keys(plan.handle, p.data_ptr(), npart, keys.data_ptr(), None), "keys")
        rt.check(rt.lib().pm_sort_by_cell(plan.handle, p.data_ptr(), npart, ks.data_ptr(),
                                          order.data_ptr(), None), "sort")
        torch.cuda.synchronize()
        want_keys = O.cell_keys(pos, cfg)
        want_order = O.sort_order(pos, cfg)
        assert np.array_equal(keys.cpu().numpy().astype(np.int64), want_keys)
        assert np.array_equal(order.cpu().numpy().astype(np.int64), want_order)
        assert np.array_equal(ks.cpu().numpy().astype(np.int64), want_keys[want_order])


# ----------------------------------------------------------------------------- deposit
@pytest.mark.parametrize("name", CASES)
def test_density_matches_golden_and_is_deterministic(pm, golden_dir, name):
    g, cfg = load_case(golden_dir, name)
    pm.set_config(cfg_ns(cfg))
    pos = dev(g["pos0"])
    rho = pm.density(pos, float(g["mass"]))
    assert rho.shape == (cfg.N_CELLS,) * 3 and rho.dtype == torch.float32 and rho.is_cuda
    got = rho.cpu().numpy()
    if "rho_0" not in g:
        g = {"rho_0": O.density(g["pos0"], float(g["mass"]), cfg), "pos0": g["pos0"], "mass": g["mass"]}
    assert rel_l2(got, g["rho_0"]) <= REL_L2
    # far tighter in practice; what is left is the REFERENCE's float32 running-sum rounding
    # (1.05e-6 on a cell holding ~1650 particles), ours accumulates in float64
    assert rel_l2(got, g["rho_0"]) <= 2e-6
    for _ in range(3):
        assert torch.equal(pm.density(pos, float(g["mass"])), rho)   # bit-reproducible
    # NumPy in -> NumPy out, same values (drop-in signature of density.py:8)
    host = pm.density(g["pos0"], float(g["mass"]))
    assert isinstance(host, np.ndarray) and np.array_equal(host, got)


def test_density_single_particle_weights_and_wrap(pm):
    cfg = O.Config(N_CELLS=16, N_PARTS=8)
    pm.set_config(cfg_ns(cfg))
    rho = pm.density(dev(np.array([[4.25], [7.0], [15.5]], dtype=np.float32)), 1.0).cpu().numpy()
    assert rho[15, 7, 4] == 0.375 and rho[15, 7, 5] == 0.125
    assert rho[0, 7, 4] == 0.375 and rho[0, 7, 5] == 0.125      # density.py:33-35
    assert rho.sum() == 1.0
    # Q4: x == N_CELLS deposits -(Nc-1)*m and +Nc*m
    rho = pm.density(dev(np.array([[16.0], [3.0], [2.0]], dtype=np.float32)), 1.0).cpu().numpy()
    assert rho[2, 3, 0] == -15.0 and rho[2, 3, 1] == 16.0
    # empty particle set -> all zeros written (no memset relied upon)
    rho = pm.density(torch.empty((3, 0), dtype=torch.float32, device="cuda"), 1.0)
    assert float(rho.abs().max()) == 0.0


@pytest.mark.parametrize("xseg", [8, 16])
def test_density_with_segmented_rows(pm, golden_dir, xseg, monkeypatch):
    """Wide meshes (1024, 2048) cut every mesh row into 512-cell segments, one warp each; force the
    same code path on the small fixtures and demand the same densities as the unsegmented kernel."""
    for name in ["free16", "clustered32", "g32_step2", "g16_free10"]:
        g, cfg = load_case(golden_dir, name)
        pm.set_config(cfg_ns(cfg))
        pos = dev(g["pos0"])
        pm.release_plans()
        monkeypatch.delenv("PM_DEPOSIT_XSEG", raising=False)
        ref = pm.density(pos, float(g["mass"]))
        pm.release_plans()
        monkeypatch.setenv("PM_DEPOSIT_XSEG", str(xseg))
        got = pm.density(pos, float(g["mass"]))
        pm.release_plans()
        assert rel_l2(got.cpu().numpy(), ref.cpu().numpy()) <= 1e-7
        want = g["rho_0"] if "rho_0" in g else O.density(g["pos0"], float(g["mass"]), cfg)
        assert rel_l2(got.cpu().numpy(), want) <= 2e-6
        # a full step through the segmented deposit (row table feeds the gather's live count too)
        p1, v1 = dev(g["pos0"]), dev(g["vel0"])
        st = pm.ResidentParticles(p1, v1)
        st.step(float(g["a_list"][0]), float(g["da"]), mass=float(g["mass"]))
        st.store(p1, v1)
        pm.release_plans()
        monkeypatch.delenv("PM_DEPOSIT_XSEG")
        p2, v2 = dev(g["pos0"]), dev(g["vel0"])
        pm.step(p2, v2, float(g["a_list"][0]), float(g["da"]), mass=float(g["mass"]))
        # the segment boundaries change the float32 summation order of the deposit (a few cells differ in
        # the last bit); on the spike fixtures one ulp of the 8192-32768-high cells is amplified to the
        # float32 floor documented at SPIKE_CASES (scratch/diag_seg.py prints the per-fixture numbers)
        assert rel_l2_periodic(p1.cpu().numpy(), p2.cpu().numpy(), cfg.N_CELLS) <= SPIKE_CASES.get(name, 1e-6)
    monkeypatch.delenv("PM_DEPOSIT_XSEG", raising=False)
    pm.release_plans()


def test_density_ragged_sizes(pm):
    # particle counts that are not multiples of 4/32 and unaligned row starts (scalar key path)
    cfg = O.Config(N_CELLS=24, N_PARTS=8)
    pm.set_config(cfg_ns(cfg))
    rs = np.random.RandomState(0)
    for npart in [1, 2, 3, 5, 31, 33, 1001]:
        pos = rs.uniform(0, 24, size=(3, npart)).astype(np.float32)
        got = pm.density(dev(pos), 2.5).cpu().numpy()
        assert rel_l2(got, O.density(pos, 2.5, cfg)) <= 2e-7


# ----------------------------------------------------------------------------- Poisson
@pytest.mark.parametrize("name", CASES)
def test_fourier_grid_and_potential(pm, golden_dir, name):
    g, cfg = load_case(golden_dir, name)
    pm.set_config(cfg_ns(cfg))
    fg = pm.fourier_grid()
    table = fg.to_array().cpu().numpy()
    want = O.fourier_grid(cfg)
    assert table.shape == want.shape and table[0, 0, 0] == 0.0
    # bit for bit: the plan holds the reference's own sin^2(k/2) values (NumPy float32, _runtime.reference_sin2_table)
    # and the table kernel adds and divides in correctly rounded float32 (fourier_utils.py:15-16)
    assert np.array_equal(table, want)
    for s in range(10):
        if f"phi_{s}" not in g:
            continue
        a = float(g["a_list"][s])
        phi = pm.potential(dev(g[f"rho_{s}"]), fg, a).cpu().numpy()
        ref = g[f"phi_{s}"].astype(np.float64)
        # DC of the Green's table is 0 on both sides; compare phi - mean anyway
        assert rel_l2(phi - phi.mean(dtype=np.float64), ref - ref.mean()) <= REL_L2


@pytest.mark.parametrize("n", [32, 64, 128, 256])
def test_hand_written_fft_against_cufft_and_float64(pm, n):
    """pm_fft.cu (default for power-of-two meshes) vs the cuFFT path vs a float64 NumPy solve."""
    cfg = O.Config(N_CELLS=n)
    pm.set_config(cfg_ns(cfg))
    rt = pm._runtime
    rs = np.random.RandomState(n)
    rho = (rs.rand(n, n, n).astype(np.float32) * 3.0)
    rho[rs.randint(n), rs.randint(n), rs.randint(n)] += 300.0       # a halo-like spike
    fg = pm.fourier_grid()
    plan = rt.get_plan(n, 1, 0)
    assert rt.lib().pm_plan_fft_backend(plan.handle) == 0           # own FFT is the default
    d = dev(rho)
    phi_own = pm.potential(d, fg, 0.37).cpu().numpy()
    assert np.array_equal(d.cpu().numpy(), rho)                     # input not modified
    rt.check(rt.lib().pm_plan_set_fft_backend(plan.handle, 1), "backend")
    phi_lib = pm.potential(d, fg, 0.37).cpu().numpy()
    rt.check(rt.lib().pm_plan_set_fft_backend(plan.handle, 0), "backend")
    assert rel_l2(phi_own, phi_lib) <= 2e-6
    if n <= 128:
        ref = O.potential(rho, O.fourier_grid(cfg), 0.37, cfg)
        assert rel_l2(phi_own, ref) <= 2e-6 and rel_l2(phi_lib, ref) <= 2e-6
    # linearity and a pure mode along each axis (exercises the packed DC/Nyquist column)
    x = np.arange(n)
    for axis in range(3):
        for m in (1, n // 2):
            shape = [1, 1, 1]
            shape[axis] = n
            mode = np.cos(2 * np.pi * m * x / n).astype(np.float32).reshape(shape)
            r = np.broadcast_to(mode, (n, n, n)).copy()
            phi = pm.potential(dev(r), fg, 0.25).cpu().numpy()
            want = -3 * cfg.OMEGA_M0 / 8 / 0.25 / np.sin(np.pi * m / n) ** 2 * r
            assert np.abs(phi - want).max() < 2e-5 * np.abs(want).max(), (axis, m)


@pytest.mark.parametrize("n", [512, 1024])
def test_hand_written_fft_large_meshes_against_cufft(pm, n):
    """BASELINE configs 2 and 3 mesh sizes: own FFT vs the cuFFT path on device-generated input."""
    cfg = O.Config(N_CELLS=n)
    pm.set_config(cfg_ns(cfg))
    rt = pm._runtime
    g = torch.Generator(device="cuda").manual_seed(n)
    rho = torch.rand((n, n, n), generator=g, device="cuda") * 4.0
    rho[n // 3, n // 5, n // 7] += 500.0
    fg = pm.fourier_grid()
    plan = rt.get_plan(n, 1, 0)
    phi_own = pm.potential(rho, fg, 0.5)
    rt.check(rt.lib().pm_plan_set_fft_backend(plan.handle, 1), "backend")
    phi_lib = pm.potential(rho, fg, 0.5)
    rt.check(rt.lib().pm_plan_set_fft_backend(plan.handle, 0), "backend")
    err = float((phi_own.double() - phi_lib.double()).norm() / phi_lib.double().norm())
    assert err <= 2e-6, err
    del phi_own, phi_lib, rho
    pm.release_plans()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n", [256, 512, 1024])
def test_fused_plane_fft_equals_separate_passes_bit_for_bit(pm, n):
    """k_fft_plane (row pass + y pass of a direction in one persistent launch, the intermediate
    plane staying in L2) does the same arithmetic as the two separate launches: the potential must
    be identical bit for bit, for the default lag, for lag 1 (consumers wait on their producers
    almost every time) and for a lag of the whole mesh (all producers first)."""
    cfg = O.Config(N_CELLS=n)
    pm.set_config(cfg_ns(cfg))
    rt = pm._runtime
    g = torch.Generator(device="cuda").manual_seed(7 * n)
    rho = torch.rand((n, n, n), generator=g, device="cuda") * 4.0
    rho[n // 3, n // 5, n // 7] += 500.0
    fg = pm.fourier_grid()
    plan = rt.get_plan(n, 1, 0)
    L = rt.lib()
    rt.check(L.pm_plan_set_fft_fuse(plan.handle, 0, 0), "fuse off")
    ref = pm.potential(rho, fg, 0.5)
    for lag in (12, 1, 3, n):
        rt.check(L.pm_plan_set_fft_fuse(plan.handle, 1, lag), "fuse on")
        for _ in range(2):
            phi = pm.potential(rho, fg, 0.5)
            assert torch.equal(phi, ref), (n, lag)
    assert L.pm_plan_fft_sync_errors(plan.handle) == 0
    rt.check(L.pm_plan_set_fft_fuse(plan.handle, 0, 12), "fuse default (off)")
    del phi, ref, rho
    pm.release_plans()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n", [256, 512, 1024])
def test_two_stage_fft_matches_three_stage_and_cufft(pm, n):
    """pm_fft2.cuh (two in-register DFT stages, natural order) against the radix-8 kernels of
    pm_fft.cu (digit-reversed order) and cuFFT: three independent implementations of the same solve."""
    cfg = O.Config(N_CELLS=n)
    pm.set_config(cfg_ns(cfg))
    rt = pm._runtime
    L = rt.lib()
    g = torch.Generator(device="cuda").manual_seed(11 * n)
    rho = torch.rand((n, n, n), generator=g, device="cuda") * 4.0
    rho[n // 2, n - 1, 0] += 700.0
    rho[0, 0, n - 1] -= 300.0
    fg = pm.fourier_grid()
    plan = rt.get_plan(n, 1, 0)
    out = {}
    for name, variant, fuse in (("v2_fused", 2, 1), ("v2_split", 2, 0), ("mix_fused", 1, 1),
                                ("mix_split", 1, 0), ("pipe2", 3, 0), ("pipe3", 4, 0), ("v1", 0, 0)):
        rt.check(L.pm_plan_set_fft_variant(plan.handle, variant), "variant")
        rt.check(L.pm_plan_set_fft_fuse(plan.handle, fuse, 0), "fuse")
        out[name] = pm.potential(rho, fg, 0.5).double()
    rt.check(L.pm_plan_set_fft_backend(plan.handle, 1), "backend")
    lib = pm.potential(rho, fg, 0.5).double()
    rt.check(L.pm_plan_set_fft_backend(plan.handle, 0), "backend")
    rt.check(L.pm_plan_set_fft_variant(plan.handle, 3), "variant (default)")
    rt.check(L.pm_plan_set_fft_fuse(plan.handle, 0, 0), "fuse")
    assert torch.equal(out["v2_fused"], out["v2_split"])
    assert torch.equal(out["mix_fused"], out["mix_split"])
    # the cp.async-pipelined y passes (pm_fft3.cuh) do the arithmetic of the two-stage kernels
    assert torch.equal(out["pipe2"], out["mix_split"])
    assert torch.equal(out["pipe3"], out["mix_split"])
    for name in ("v2_fused", "mix_split", "v1"):
        err = float((out[name] - lib).norm() / lib.norm())
        assert err <= 2e-6, (name, err)
    assert L.pm_plan_fft_sync_errors(plan.handle) == 0
    del out, lib, rho
    pm.release_plans()
    torch.cuda.empty_cache()


def test_single_mode_potential(pm):
    cfg = O.Config(N_CELLS=32)
    pm.set_config(cfg_ns(cfg))
    m, a = 3, 0.25
    x = np.arange(32)
    rho = np.broadcast_to(np.cos(2 * np.pi * m * x / 32).astype(np.float32), (32, 32, 32)).copy()
    phi = pm.potential(dev(rho), pm.fourier_grid(), a).cpu().numpy()
    want = -3 * cfg.OMEGA_M0 / 8 / a / np.sin(np.pi * m / 32) ** 2 * rho
    assert np.abs(phi - want).max() < 2e-5 * np.abs(want).max()


# ----------------------------------------------------------------------------- gather/kick/drift
@pytest.mark.parametrize("name", CASES)
def test_integrate_bit_exact_given_reference_phi(pm, golden_dir, name):
    g, cfg = load_case(golden_dir, name)
    pm.set_config(cfg_ns(cfg))
    da = float(g["da"])
    s = 0
    if "phi_0" not in g:
        pytest.skip("fixture keeps no step-0 potential")
    a = float(g["a_list"][s])
    f_a1 = float(O.f(a + da, [cfg.H0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0]))
    pos, vel = dev(g["pos0"]), dev(g["vel0"])
    p2, v2 = pm.integrate(pos, vel, a, f_a1, da, dev(g[f"phi_{s}"]))
    assert p2 is pos and v2 is vel
    assert np.array_equal(vel.cpu().numpy(), g["vel_1"])
    assert np.array_equal(pos.cpu().numpy(), g["pos_1"])
    # accelerations (g_p of integrate.py:92) against the oracle
    npart = g["pos0"].shape[1]
    acc_ref = np.zeros((3, npart))
    po, vo = g["pos0"].copy(), g["vel0"].copy()
    O.integrate(po, vo, a, f_a1, da, g[f"phi_{s}"], cfg, acc=acc_ref)
    acc = torch.zeros((3, npart), dtype=torch.float32, device="cuda")
    import importlib
    mod = importlib.import_module("cosmological_particle_mesh_simulation_b200.integrate")
    mod._integrate_device(dev(g["pos0"]), dev(g["vel0"]), a, f_a1, da, dev(g[f"phi_{s}"]), acc=acc)
    assert np.array_equal(acc.cpu().numpy(), acc_ref.astype(np.float32))


def test_uniform_lattice_zero_kick(pm):
    cfg = O.Config(N_CELLS=16, N_PARTS=16)
    pm.set_config(cfg_ns(cfg))
    pos, vel = O.lattice_ic(16, 16, jitter=0.0)
    p, v = dev(pos), dev(vel)
    rho = pm.density(p, 1.0)
    assert torch.all(rho == 1.0)
    pm.advance_time(rho, p, v, pm.fourier_grid(), 0.5, 0.01)
    assert float(v.abs().max()) <= 1e-6     # phi is constant up to float32 FFT rounding


# ----------------------------------------------------------------------------- whole step
@pytest.mark.parametrize("name", CASES)
def test_free_running_steps_against_golden(pm, golden_dir, name):
    g, cfg = load_case(golden_dir, name)
    pm.set_config(cfg_ns(cfg))
    n = cfg.N_CELLS
    tol = SPIKE_CASES.get(name, REL_L2)
    pos, vel = dev(g["pos0"]), dev(g["vel0"])
    fg = pm.fourier_grid()
    da = float(g["da"])
    for s, a in enumerate(g["a_list"]):
        if name in SPIKE_CASES and s >= 2:
            break   # beyond two steps a stress fixture only measures its own chaos (Q4 on/off flips)
        rho = pm.density(pos, float(g["mass"]))
        if f"rho_{s}" in g:
            assert rel_l2(rho.cpu().numpy(), g[f"rho_{s}"]) <= tol, f"density step {s}"
        pm.advance_time(rho, pos, vel, fg, float(a), da)
        if f"pos_{s + 1}" in g:
            assert rel_l2_periodic(pos.cpu().numpy(), g[f"pos_{s + 1}"], n) <= tol, f"positions step {s}"
            assert rel_l2(vel.cpu().numpy(), g[f"vel_{s + 1}"]) <= tol, f"velocities step {s}"


@pytest.mark.parametrize("name", sorted(SPIKE_CASES))
def test_spike_fixtures_meet_1e_5_when_only_the_transform_precision_changes(pm, golden_dir, name, monkeypatch):
    """The three fixtures whose float32-transform tolerance is relaxed (SPIKE_CASES) against the north_star bar,
    ALL their steps, with one thing changed: the Poisson solve transforms in float64 (PM_FFT_BACKEND=f64 ->
    cuFFT D2Z/Z2D, pm_poisson.cu), as the reference's complex128 pyFFTW does (src/potential.py:19-29).
    Deposit, Green's table (the reference's own sin^2 values, exact float32 reciprocal), gather, kick and
    drift are the production kernels.  Measured: positions <= 5e-8, velocities <= 8e-8, density <= 9e-7."""
    monkeypatch.setenv("PM_FFT_BACKEND", "f64")
    pm.release_plans()
    try:
        g, cfg = load_case(golden_dir, name)
        pm.set_config(cfg_ns(cfg))
        n = cfg.N_CELLS
        pos, vel = dev(g["pos0"]), dev(g["vel0"])
        fg = pm.fourier_grid()
        da = float(g["da"])
        checked = 0
        for s, a in enumerate(g["a_list"]):
            rho = pm.density(pos, float(g["mass"]))
            if f"rho_{s}" in g:
                assert rel_l2(rho.cpu().numpy(), g[f"rho_{s}"]) <= REL_L2, f"density step {s}"
            pm.advance_time(rho, pos, vel, fg, float(a), da)
            if f"pos_{s + 1}" in g:
                assert rel_l2_periodic(pos.cpu().numpy(), g[f"pos_{s + 1}"], n) <= REL_L2, f"positions step {s}"
                assert rel_l2(vel.cpu().numpy(), g[f"vel_{s + 1}"]) <= REL_L2, f"velocities step {s}"
                checked += 1
        assert checked >= 2
        from cosmological_particle_mesh_simulation_b200 import _runtime as rt
        plan = rt.get_plan(n, pos.shape[1], pos.device.index)
        assert rt.lib().pm_plan_fft_backend(plan.handle) == 2
    finally:
        pm.release_plans()          # the next test builds its plans without the override


@pytest.mark.parametrize("name", ["free16", "free32", "clustered32", "g12_nonpow2"])
def test_resident_state_matches_stateless_steps_and_golden(pm, golden_dir, name):
    """load -> n x pm_step_resident -> store against n x pm_step and the golden particles; the
    storage order after each step is the stable cell sort of the previous storage order."""
    g, cfg = load_case(golden_dir, name)
    pm.set_config(cfg_ns(cfg))
    n, mass, da = cfg.N_CELLS, float(g["mass"]), float(g["da"])
    tol = SPIKE_CASES.get(name, REL_L2)
    pos, vel = dev(g["pos0"]), dev(g["vel0"])
    ps, vs = pos.clone(), vel.clone()
    state = pm.ResidentParticles(pos, vel)
    out_p, out_v = torch.empty_like(pos), torch.empty_like(vel)
    npart = pos.shape[1]
    prev_ids = np.arange(npart)
    prev_pos = g["pos0"]
    for s, a in enumerate(g["a_list"]):
        rho_r, rho_s = torch.empty((n, n, n), device="cuda"), torch.empty((n, n, n), device="cuda")
        state.step(float(a), da, mass=mass, rho_out=rho_r)
        pm.step(ps, vs, float(a), da, mass=mass, rho_out=rho_s)
        state.store(out_p, out_v)
        # same particles, original order; summation order inside a cell may differ (ties follow
        # the storage order), hence a rounding-level tolerance instead of equality
        assert rel_l2(rho_r.cpu().numpy(), rho_s.cpu().numpy()) <= 1e-6
        assert rel_l2_periodic(out_p.cpu().numpy(), ps.cpu().numpy(), n) <= 1e-6
        assert rel_l2(out_v.cpu().numpy(), vs.cpu().numpy()) <= 1e-6
        if f"pos_{s + 1}" in g:
            assert rel_l2_periodic(out_p.cpu().numpy(), g[f"pos_{s + 1}"], n) <= tol
            assert rel_l2(out_v.cpu().numpy(), g[f"vel_{s + 1}"]) <= tol
        # sort order, bit-exact: new storage order = previous order stably re-sorted by the cell
        # keys of the positions the step started from
        ids = state.order().cpu().numpy().astype(np.int64)
        assert np.array_equal(np.sort(ids), np.arange(npart))
        keys_prev = O.cell_keys(np.ascontiguousarray(prev_pos), cfg)
        want = prev_ids[np.argsort(keys_prev[prev_ids], kind="stable")]
        assert np.array_equal(ids, want)
        prev_ids, prev_pos = ids, out_p.cpu().numpy()


@pytest.mark.parametrize("n_parts,n_cells,vsig", [(32, 64, 0.05), (48, 64, 0.3), (40, 128, 0.02)])
def test_incremental_sort_equals_full_sort_bit_for_bit(pm, n_parts, n_cells, vsig):
    """The resident path re-sorts only the particles whose cell changed and merges them into the
    still-sorted rest (pm_sort.cu).  Storage order, positions and velocities must equal those of the
    full radix sort bit for bit, step after step, with a realistic and with a large mover fraction
    (the latter crosses the 40 % threshold and exercises the fall-back)."""
    cfg = types.SimpleNamespace(N_CELLS=n_cells, N_PARTS=n_parts, OMEGA_M0=0.31, OMEGA_K0=0.0,
                                OMEGA_LAMBDA0=0.69, H0=0.68, A_INIT=0.01, A_END=1.0, STEPS=1000)
    pm.set_config(cfg)
    rng = np.random.default_rng(5)
    npart = n_parts ** 3 - 37           # ragged: not a multiple of the 2048-entry tile or of 4
    pos = rng.uniform(0, n_cells, (3, npart)).astype(np.float32)
    pos[:, : npart // 8] = (n_cells / 2 + rng.normal(0, 1.5, (3, npart // 8))).astype(np.float32) % n_cells
    vel = rng.normal(0, vsig, (3, npart)).astype(np.float32)
    sched = pm.loop_scale_factors(cfg)[:6]
    out = {}
    for mode in ("full", "auto"):
        st = pm.ResidentParticles(dev(pos), dev(vel))
        st.set_sort_mode(mode)
        seen = []
        for a, da in sched:
            st.step(a, da)
            torch.cuda.synchronize()
            seen.append(st.sort_stats())
        p, v = torch.empty(3, npart, device="cuda"), torch.empty(3, npart, device="cuda")
        st.store(p, v)
        out[mode] = (st.order().cpu().numpy(), p.cpu().numpy(), v.cpu().numpy(), seen)
        st.set_sort_mode("auto")
    assert np.array_equal(out["full"][0], out["auto"][0])
    assert np.array_equal(out["full"][1], out["auto"][1])
    assert np.array_equal(out["full"][2], out["auto"][2])
    assert all(s[2] == "full" for s in out["full"][3])
    modes = [s[2] for s in out["auto"][3]]
    assert modes[0] == "full"                                   # no previous order on the first step
    fracs = [s[1] / s[0] for s in out["auto"][3][1:]]
    if vsig <= 0.05:
        assert all(m == "incremental" for m in modes[1:]), modes
        assert all(0.0 < f < 0.4 for f in fracs), fracs
    assert sorted(out["auto"][0].tolist()) == list(range(npart))


@pytest.mark.parametrize("n_parts,n_cells", [(64, 128), (96, 256)])
def test_graph_replay_equals_eager_steps_bit_for_bit(pm, n_parts, n_cells):
    """The steady-state resident step replays as a CUDA graph when it runs on a non-default stream
    (pm_api.cu, resident_step_graphed): two graphs, one per buffer-set parity, re-parameterised through
    one kernel node per step.  Eight steps (capture x2, then replays of both parities with NEW scalars
    every step) must leave the same storage order, positions, velocities and density as eager steps."""
    import ctypes
    cfg = types.SimpleNamespace(N_CELLS=n_cells, N_PARTS=n_parts, OMEGA_M0=0.31, OMEGA_K0=0.0,
                                OMEGA_LAMBDA0=0.69, H0=0.68, A_INIT=0.01, A_END=1.0, STEPS=1000)
    pm.set_config(cfg)
    rt = pm._runtime
    rng = np.random.default_rng(11)
    npart = n_parts ** 3 - 5
    pos = rng.uniform(0, n_cells, (3, npart)).astype(np.float32)
    vel = rng.normal(0, 0.05, (3, npart)).astype(np.float32)
    sched = pm.loop_scale_factors(cfg)[:8]
    out = {}
    for mode in ("eager", "graph"):
        st = pm.ResidentParticles(dev(pos), dev(vel))
        rt.check(rt.lib().pm_plan_set_graph(st.plan.handle, 1 if mode == "graph" else 0), "pm_plan_set_graph")
        rho = torch.zeros((n_cells,) * 3, device="cuda")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for a, da in sched:
                st.step(a, da, rho_out=rho)
        side.synchronize()
        replays = int(rt.lib().pm_plan_graph_replays(st.plan.handle))
        p, v = torch.empty(3, npart, device="cuda"), torch.empty(3, npart, device="cuda")
        st.store(p, v)
        out[mode] = (st.order().cpu().numpy(), p.cpu().numpy(), v.cpu().numpy(), rho.cpu().numpy(), replays, st.sort_stats())
        st.close()
    assert out["eager"][4] == 0
    assert out["graph"][4] == len(sched) - 1, "every step after the first (full sort, not steady state) is a replay"
    for k in range(4):
        assert np.array_equal(out["eager"][k], out["graph"][k]), ("order", "positions", "velocities", "density")[k]


@pytest.mark.parametrize("name", ["g16_free10", "clustered32"])
def test_fused_step_and_host_step_equal_the_composed_calls(pm, golden_dir, name):
    g, cfg = load_case(golden_dir, name)
    pm.set_config(cfg_ns(cfg))
    da, a, mass = float(g["da"]), float(g["a_list"][0]), float(g["mass"])
    pos, vel = dev(g["pos0"]), dev(g["vel0"])
    rho = pm.density(pos, mass)
    pm.advance_time(rho, pos, vel, pm.fourier_grid(), a, da)
    p2, v2 = dev(g["pos0"]), dev(g["vel0"])
    rho2 = torch.empty_like(rho)
    pm.step(p2, v2, a, da, mass=mass, rho_out=rho2)
    assert torch.equal(rho2, rho) and torch.equal(p2, pos) and torch.equal(v2, vel)
    p3, v3 = dev(g["pos0"]), dev(g["vel0"])
    pm.step(p3, v3, a, da, mass=mass)                      # density kept in the plan's mesh
    assert torch.equal(p3, pos) and torch.equal(v3, vel)
    ph = torch.from_numpy(g["pos0"].copy()).pin_memory()
    vh = torch.from_numpy(g["vel0"].copy()).pin_memory()
    rh = torch.empty(rho.shape, dtype=torch.float32).pin_memory()
    pm.step_host(ph, vh, a, da, mass=mass, rho_out=rh)
    assert torch.equal(ph, pos.cpu()) and torch.equal(vh, vel.cpu()) and torch.equal(rh, rho.cpu())
    pn, vn = g["pos0"].copy(), g["vel0"].copy()            # pageable NumPy buffers
    pm.step_host(pn, vn, a, da, mass=mass)
    assert np.array_equal(pn, pos.cpu().numpy()) and np.array_equal(vn, vel.cpu().numpy())


@pytest.mark.parametrize("n_cells,n_parts,blob", [(128, 64, 0), (256, 96, 0), (128, 50, 40000), (512, 128, 0)])
def test_host_step_split_gather_equals_the_device_step_bit_for_bit(pm, n_cells, n_parts, blob):
    """pm_step_host on meshes that have the warp-specialised gather (128 / 256 / 512) splits the step where the
    velocities enter it: stencil sums of the cell-ordered particles stored at the original index (k_gather_ws,
    SONLY), then kick + drift in the caller's order behind the chunked velocity upload (k_push_rows).  Same
    pm_push, so the results must equal pm.step's on the same arrays bit for bit -- incl. a crowded blob (work
    list), particle counts that are not multiples of the chunk granularity, and positions == N_CELLS (Q4)."""
    cfg = O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100)
    pm.set_config(cfg_ns(cfg))
    rng = np.random.default_rng(n_cells + n_parts)
    npart = n_parts ** 3
    pos_h = rng.uniform(0, n_cells, (3, npart)).astype(np.float32)
    if blob:
        pos_h[:, :blob] = (n_cells / 2 + rng.normal(0, 1.0, (3, blob))).astype(np.float32) % n_cells
    pos_h[:, -1] = [float(n_cells), 0.0, n_cells - 1e-3]
    pos_h[:, -2] = [3.5, float(n_cells), float(n_cells)]
    vel_h = rng.normal(0, 0.5, (3, npart)).astype(np.float32)
    mass = (n_cells / n_parts) ** 3
    a, da = 0.3, 0.0099
    pd, vd = dev(pos_h), dev(vel_h)
    rho_d = torch.empty((n_cells,) * 3, dtype=torch.float32, device="cuda")
    ph, vh = torch.from_numpy(pos_h.copy()).pin_memory(), torch.from_numpy(vel_h.copy()).pin_memory()
    rh = torch.empty((n_cells,) * 3, dtype=torch.float32).pin_memory()
    for k in range(3):                       # three consecutive steps: the output of one is the input of the next
        pm.step(pd, vd, a + k * da, da, mass=mass, rho_out=rho_d)
        pm.step_host(ph, vh, a + k * da, da, mass=mass, rho_out=rh)
        assert torch.equal(ph, pd.cpu()), f"positions, step {k}"
        assert torch.equal(vh, vd.cpu()), f"velocities, step {k}"
        assert torch.equal(rh, rho_d.cpu()), f"density, step {k}"
    pn, vn = pos_h.copy(), vel_h.copy()      # pageable NumPy buffers, no density
    pm.step_host(pn, vn, a, da, mass=mass)
    p1, v1 = dev(pos_h), dev(vel_h)
    pm.step(p1, v1, a, da, mass=mass)
    assert np.array_equal(pn, p1.cpu().numpy()) and np.array_equal(vn, v1.cpu().numpy())


def test_numpy_drop_in_signatures(pm, golden_dir):
    """The reference's loop body (pmesh.py:60-61) on NumPy arrays, unchanged call shapes."""
    g, cfg = load_case(golden_dir, "free16")
    pm.set_config(cfg_ns(cfg))
    pos, vel = g["pos0"].copy(), g["vel0"].copy()
    fg = pm.fourier_grid()
    rho = pm.density(pos, float(g["mass"]))
    p2, v2 = pm.advance_time(rho, pos, vel, fg, float(g["a_list"][0]), float(g["da"]))
    assert p2 is pos and v2 is vel and isinstance(rho, np.ndarray)
    assert rel_l2_periodic(pos, g["pos_1"], cfg.N_CELLS) <= REL_L2
    assert rel_l2(vel, g["vel_1"]) <= REL_L2


def test_numpy_drop_in_loop_pins_the_callers_arrays_and_pools_the_results(pm):
    """The reference's loop body on NumPy arrays (src/pmesh.py:60-61), several steps: the two particle arrays are
    page-locked in place once, density() hands out pooled pinned result arrays and never one that is still
    referenced (directly or through a view), and the numbers are those of the same calls on CUDA tensors."""
    from cosmological_particle_mesh_simulation_b200 import _runtime as rt
    cfg = O.Config(N_CELLS=128, N_PARTS=64, STEPS=100)
    pm.set_config(cfg_ns(cfg))
    pos_h, vel_h = O.lattice_ic(64, 128, seed=5, vel_rms=0.3)
    pos, vel = pos_h.copy(), vel_h.copy()
    pd, vd = dev(pos_h), dev(vel_h)
    fg = pm.fourier_grid()
    pm.set_resident_dropin(False)              # the CUDA-tensor side runs the same stateless kernels
    try:
        kept, a, da = [], 0.3, 0.0099
        for k in range(4):
            rho = pm.density(pos, 8.0)                                  # pmesh.py:60
            rho_d = pm.density(pd, 8.0)
            assert isinstance(rho, np.ndarray) and rho.dtype == np.float32 and rho.shape == (128,) * 3
            assert np.array_equal(rho, rho_d.cpu().numpy())
            kept.append((rho[3], rho[3].copy()))                         # a VIEW kept by the caller, and its contents
            pos, vel = pm.advance_time(rho, pos, vel, fg, a, da)        # pmesh.py:61
            pd, vd = pm.advance_time(rho_d, pd, vd, fg, a, da)
            assert np.array_equal(pos, pd.cpu().numpy()) and np.array_equal(vel, vd.cpu().numpy())
            a += da
        for view, want in kept:                                          # no pooled buffer was reused under a live view
            assert np.array_equal(view, want)
        key = (pos.ctypes.data, pos.nbytes)
        if rt._pin_enabled:
            assert key in rt._pinned_ranges and (vel.ctypes.data, vel.nbytes) in rt._pinned_ranges
            assert 1 <= len(rt._host_pool) <= rt._HOST_POOL_MAX
        del pos, vel, kept, rho, view
        import gc
        gc.collect()
        assert key not in rt._pinned_ranges                              # un-registered with the array
    finally:
        pm.set_resident_dropin(True)


def test_lazy_drop_in_equals_eager_drop_in_bit_for_bit(pm, golden_dir):
    """set_resident_dropin("lazy"): the reference's loop body (src/pmesh.py:60-61, names rebound to the
    returned objects) runs on the resident state with NO per-step write-back; the caller-visible tensors
    are brought up to date when the handles are first used.  Same bits as the default (eager) session."""
    from cosmological_particle_mesh_simulation_b200 import _session as S
    g, cfg = load_case(golden_dir, "free32")
    pm.set_config(cfg_ns(cfg))
    n, mass, da = cfg.N_CELLS, float(g["mass"]), float(g["da"])
    fg = pm.fourier_grid()
    out = {}
    try:
        for mode in (True, "lazy"):
            pm.forget_resident()
            pm.set_resident_dropin(mode)
            positions, velocities = dev(g["pos0"]), dev(g["vel0"])
            p0 = positions
            rhos = []
            for a in g["a_list"][:5]:
                rho = pm.density(positions, mass)                                                  # pmesh.py:60
                positions, velocities = pm.advance_time(rho, positions, velocities, fg, float(a), da)   # pmesh.py:61
                rhos.append(rho)
            if mode == "lazy":
                assert type(positions) is S.ResidentView and S._session.pending
                assert tuple(positions.shape) == (3, p0.shape[1]) and S._session.pending        # metadata: still deferred
                stale = p0.cpu().numpy()                 # the ORIGINAL object, read behind the library's back
                assert np.array_equal(stale, g["pos0"])  # never written so far (documented hazard)
            out[mode] = (positions.cpu().numpy(), velocities.cpu().numpy(), rhos[-1].cpu().numpy())
            if mode == "lazy":
                assert not S._session.pending and np.array_equal(p0.cpu().numpy(), out[mode][0])   # same storage, now current
                # the session goes on after a read; an in-place change through the handle ends it
                rho = pm.density(positions, mass)
                assert S._session.matches_positions(S.unwrap(positions), n)
                positions += 0.0
                assert not S._session.matches_positions(S.unwrap(positions), n)
    finally:
        pm.forget_resident()
        pm.set_resident_dropin(True)
    for k in range(3):
        assert np.array_equal(out[True][k], out["lazy"][k])
    assert rel_l2_periodic(out["lazy"][0], g["pos_5"], n) <= REL_L2 and rel_l2(out["lazy"][1], g["vel_5"]) <= REL_L2


@pytest.mark.parametrize("mode", [True, "lazy"])
def test_drop_in_session_ends_when_the_package_itself_overwrites_the_tensors(pm, golden_dir, mode):
    """The package's own entry points write through raw pointers, which torch's version counters do not see
    by themselves: pm.step / integrate / ResidentParticles.store on the tensors a drop-in session mirrors must
    end that session (and, in lazy mode, first bring the caller's bytes up to date), or the next density() /
    advance_time() would continue from a resident state the caller has just overwritten.  Checked against the
    stateless calls."""
    from cosmological_particle_mesh_simulation_b200 import _session as S
    g, cfg = load_case(golden_dir, "free32")
    pm.set_config(cfg_ns(cfg))
    n, mass, da = cfg.N_CELLS, float(g["mass"]), float(g["da"])
    a0, a1, a2, a3 = (float(x) for x in g["a_list"][:4])
    fg = pm.fourier_grid()

    def sequence():
        positions, velocities = dev(g["pos0"]), dev(g["vel0"])
        rho = pm.density(positions, mass)
        positions, velocities = pm.advance_time(rho, positions, velocities, fg, a0, da)        # a session is born here
        pm.step(positions, velocities, a1, da, mass=mass)                                      # raw write no. 1
        rho = pm.density(positions, mass)
        positions, velocities = pm.advance_time(rho, positions, velocities, fg, a2, da)
        phi = pm.potential(pm.density(positions, mass), fg, a3)
        pm.integrate(positions, velocities, a3, float(pm.f(a3 + da, [cfg.H0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0])), da, phi)   # no. 2
        rho = pm.density(positions, mass)
        other = pm.ResidentParticles(dev(g["pos0"]), dev(g["vel0"]))
        positions, velocities = pm.advance_time(rho, positions, velocities, fg, a3, da)
        other.store(positions, velocities)                                                     # no. 3: back to the start
        other.close()
        rho_end = pm.density(positions, mass)
        return positions.cpu().numpy(), velocities.cpu().numpy(), rho.cpu().numpy(), rho_end.cpu().numpy()

    try:
        pm.forget_resident()
        pm.set_resident_dropin(False)
        want = sequence()
        pm.set_resident_dropin(mode)
        got = sequence()
        assert S._session is not None
    finally:
        pm.forget_resident()
        pm.set_resident_dropin(True)
    # writes no. 1 and 2: the density after step / advance / integrate.  (Not bitwise: the stateless Poisson call
    # measures the mesh mean, the session knows it analytically -- float32 transform noise apart.  A session that
    # had gone on from its stale state would be off by a whole step, ~1e-3.)
    assert rel_l2(got[2], want[2]) <= 5e-6
    # write no. 3 put the initial state back: everything after it is exact
    for k, name in ((0, "positions"), (1, "velocities"), (3, "final density")):
        assert np.array_equal(got[k], want[k]), name
    assert np.array_equal(got[0], g["pos0"]) and np.array_equal(got[1], g["vel0"])


def test_plans_on_two_devices_in_one_process(pm):
    """One process, two GPUs: the opt-ins for > 48 KB of dynamic shared memory (FFT passes, tile deposit, gather)
    are per DEVICE; a process-wide "already set" flag would leave the second device without them and every
    launch there would fail (round-1 advisor finding).  The same resident run on cuda:0 and cuda:1, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = types.SimpleNamespace(N_CELLS=128, N_PARTS=64, OMEGA_M0=0.31, OMEGA_K0=0.0,
                                OMEGA_LAMBDA0=0.69, H0=0.68, A_INIT=0.01, A_END=1.0, STEPS=1000)
    pm.set_config(cfg)
    rng = np.random.default_rng(3)
    npart = 64 ** 3
    pos = rng.uniform(0, 128, (3, npart)).astype(np.float32)
    vel = rng.normal(0, 0.05, (3, npart)).astype(np.float32)
    sched = pm.loop_scale_factors(cfg)[:4]
    out = []
    for d in (0, 1):
        p, v = torch.from_numpy(pos).to(f"cuda:{d}"), torch.from_numpy(vel).to(f"cuda:{d}")
        st = pm.ResidentParticles(p, v)
        rho = torch.zeros((128,) * 3, device=f"cuda:{d}")
        for a, da in sched:
            st.step(a, da, rho_out=rho)
        st.store(p, v)
        out.append((p.cpu().numpy(), v.cpu().numpy(), rho.cpu().numpy()))
        st.close()
        # the stateless entry points (shared cached plan of that device) as well
        rho2 = pm.density(torch.from_numpy(pos).to(f"cuda:{d}"), 8.0)
        assert rho2.device.index == d and np.isfinite(float(rho2.sum()))
    for k in range(3):
        assert np.array_equal(out[0][k], out[1][k])


@pytest.mark.parametrize("n", [32, 20])
def test_poisson_options_deconvolution_and_spectral_gradient(pm, n):
    """pm_plan_set_poisson_options (BASELINE north_star (2); SURVEY Q6: options the reference does not have, off
    in parity mode).  Off: the solve is the default one, bit for bit.  deconvolve = p: phi_k / W(k)^p against a
    float64 NumPy evaluation.  kspace_gradient: the accelerations of the fused gather equal the CIC
    interpolation of irfftn(-i k phi_k) evaluated in float64 (relative L2 <= 1e-5)."""
    cfg = O.Config(N_CELLS=n, N_PARTS=n // 2, STEPS=100)
    pm.set_config(cfg_ns(cfg))
    rng = np.random.default_rng(7)
    rho_h = (1.0 + 0.3 * rng.standard_normal((n, n, n))).astype(np.float32)
    a = 0.3
    fg = pm.fourier_grid()
    G = O.fourier_grid(cfg).astype(np.float64)
    G[0, 0, 0] = 0.0
    k1 = 2 * np.pi * np.fft.fftfreq(n)
    w1 = np.ones(n)
    w1[1:] = (np.sin(k1[1:] / 2) / (k1[1:] / 2)) ** 2
    W = w1[:, None, None] * w1[None, :, None] * w1[None, None, :]
    rho_k = np.fft.fftn(rho_h.astype(np.float64))
    try:
        pm.release_plans()
        assert pm.poisson_options() == (0, 0)
        phi0 = pm.potential(dev(rho_h), fg, a).cpu().numpy()
        # ---- deconvolution ----
        for p in (1, 2):
            pm.set_poisson_options(deconvolve=p)
            phi = pm.potential(dev(rho_h), fg, a).cpu().numpy()
            want = np.fft.ifftn(-3 * cfg.OMEGA_M0 / 8 / a * G * rho_k / W ** p).real
            assert rel_l2(phi - phi.mean(dtype=np.float64), want - want.mean()) <= REL_L2, p
        # ---- spectral gradient (with deconvolve = 2, the usual pairing) ----
        pm.set_poisson_options(deconvolve=2, kspace_gradient=True)
        phi = pm.potential(dev(rho_h), fg, a)             # also fills the plan's three force meshes
        phi_k = -3 * cfg.OMEGA_M0 / 8 / a * G * rho_k / W ** 2
        kk = k1.copy()
        if n % 2 == 0:
            kk[n // 2] = 0.0                              # the Nyquist mode has no derivative
        F = [np.fft.ifftn(-1j * kk.reshape([-1 if ax == d else 1 for ax in range(3)]) * phi_k).real for d in (2, 1, 0)]  # x, y, z <-> axes 2, 1, 0
        npart = 5000
        pos_h = rng.uniform(0, n, (3, npart)).astype(np.float32)
        vel_h = np.zeros((3, npart), np.float32)
        acc = torch.zeros((3, npart), device="cuda")
        pos, vel = dev(pos_h.copy()), dev(vel_h.copy())
        da, fa1 = 0.0099, 1.3
        from cosmological_particle_mesh_simulation_b200.integrate import _integrate_device
        _integrate_device(pos, vel, a, fa1, da, phi, acc=acc)
        c = np.floor(pos_h).astype(np.int64) % n
        d = pos_h.astype(np.float64) - c
        want_acc = np.zeros((3, npart))
        for oz in (0, 1):
            for oy in (0, 1):
                for ox in (0, 1):
                    wgt = (d[0] if ox else 1 - d[0]) * (d[1] if oy else 1 - d[1]) * (d[2] if oz else 1 - d[2])
                    iz, iy, ix = (c[2] + oz) % n, (c[1] + oy) % n, (c[0] + ox) % n
                    for k in range(3):
                        want_acc[k] += wgt * F[k][iz, iy, ix]
        got = acc.cpu().numpy()
        assert rel_l2(got, want_acc) <= REL_L2
        assert rel_l2(vel.cpu().numpy(), da * fa1 * want_acc) <= REL_L2           # the kick used them
        # ---- and off again: the default solve, bit for bit ----
        pm.set_poisson_options()
        assert np.array_equal(pm.potential(dev(rho_h), fg, a).cpu().numpy(), phi0)
    finally:
        pm.set_poisson_options()
        pm.release_plans()


@pytest.mark.parametrize("n", [32, 64, 128])
def test_device_power_spectrum_matches_estimator(pm, n):
    """pm_power_spectrum (forward half of the hand-written FFT + on-device binning) against the
    NumPy estimator the parity tests use, on a clustered density."""
    cfg = O.Config(N_CELLS=n, N_PARTS=n // 2)
    pm.set_config(cfg_ns(cfg))
    rs = np.random.RandomState(n)
    npart = (n // 2) ** 3
    blob = rs.normal(n / 2, n / 10.0, size=(3, npart // 2))
    uni = rs.uniform(0, n, size=(3, npart - npart // 2))
    pos = (np.concatenate([blob, uni], axis=1) % n).astype(np.float32)
    rho = pm.density(dev(np.ascontiguousarray(pos)), 8.0)
    k, p = pm.analysis.power_spectrum(rho)
    kc, pc = O.power_spectrum(rho.cpu().numpy())
    assert k.shape[0] == n // 2 - 1 and len(pc) == n // 2 - 1
    assert np.allclose(p.cpu().numpy(), pc, rtol=2e-5, atol=0)
    proj = pm.analysis.project(rho, 5).cpu().numpy()
    assert np.allclose(proj, rho.cpu().numpy()[:5].astype(np.float64).sum(axis=0), rtol=1e-12)


def test_simulator_loop_equals_manual_steps(pm, golden_dir):
    """pmesh.simulator: the reference's while-loop (pmesh.py:56-63) on resident state; on_step sees
    the PRE-step density with the POST-step particles."""
    g, cfg = load_case(golden_dir, "free16")
    c = cfg_ns(cfg)
    pm.set_config(c)
    pos, vel = dev(g["pos0"]), dev(g["vel0"])
    seen = []

    def on_step(i, a, rho, p, v):
        seen.append((i, a, float(rho.sum(dtype=torch.float64)), p.clone(), v.clone()))

    pm.simulator(pos, vel, on_step=on_step, max_steps=3)
    p2, v2 = dev(g["pos0"]), dev(g["vel0"])
    sched = pm.loop_scale_factors(c)
    for i in range(3):
        a, da = sched[i]
        rho = pm.density(p2, (c.N_CELLS / c.N_PARTS) ** 3)
        pm.advance_time(rho, p2, v2, pm.fourier_grid(), a, da)
        assert seen[i][0] == i and seen[i][1] == a + da
        assert abs(seen[i][2] - float(rho.sum(dtype=torch.float64))) <= 1e-6 * abs(seen[i][2])
        assert rel_l2_periodic(seen[i][3].cpu().numpy(), p2.cpu().numpy(), c.N_CELLS) <= 1e-6
    assert rel_l2_periodic(pos.cpu().numpy(), g["pos_3"], c.N_CELLS) <= REL_L2
    assert rel_l2(vel.cpu().numpy(), g["vel_3"]) <= REL_L2


def test_errors_are_loud(pm):
    cfg = O.Config(N_CELLS=16, N_PARTS=8)
    pm.set_config(cfg_ns(cfg))
    with pytest.raises(TypeError):
        pm.density(torch.zeros((3, 8), dtype=torch.float64, device="cuda"), 1.0)
    with pytest.raises(TypeError):
        pm.potential(torch.zeros((16, 16, 16), device="cuda"), np.zeros((16, 16, 16), np.float32), 0.5)
    with pytest.raises(pm.PMStepError):
        pm._runtime.Plan(2000, 10, 0)       # n_cells^3 >= 2^32 is outside this build


# ----------------------------------------------------------------------------- full-size properties
def test_full_size_properties_256_on_512(pm):
    """BASELINE config 2 (256^3 on 512^3): too big for the oracle in a test, so check
    size-independent properties: mass conservation, run-to-run determinism, sortedness,
    momentum-free uniform lattice, and every particle inside [0, Nc] after a step."""
    cfg = O.Config(N_CELLS=512, N_PARTS=256)
    pm.set_config(cfg_ns(cfg))
    rt = pm._runtime
    pos_h, vel_h = O.lattice_ic(256, 512, seed=38)
    npart = pos_h.shape[1]
    pos, vel = dev(pos_h), dev(vel_h)
    rho = pm.density(pos, 8.0)
    total = float(rho.sum(dtype=torch.float64))
    assert abs(total - 8.0 * npart) <= 1e-6 * 8.0 * npart
    assert torch.equal(pm.density(pos, 8.0), rho)
    plan = rt.get_plan(512, npart, 0)
    ks = torch.empty(npart, dtype=torch.int32, device="cuda")
    order = torch.empty_like(ks)
    rt.check(rt.lib().pm_sort_by_cell(plan.handle, pos.data_ptr(), npart, ks.data_ptr(),
                                      order.data_ptr(), None), "sort")
    assert bool((ks[1:] >= ks[:-1]).all())
    assert torch.equal(torch.sort(order.long()).values, torch.arange(npart, device="cuda"))
    # spot-check 100k random particles' keys and a z-slab of the density against the oracle
    sel = np.random.RandomState(1).choice(npart, 100000, replace=False)
    keys = torch.empty(npart, dtype=torch.int32, device="cuda")
    rt.check(rt.lib().pm_cell_keys(plan.handle, pos.data_ptr(), npart, keys.data_ptr(), None), "keys")
    assert np.array_equal(keys.cpu().numpy()[sel].astype(np.int64),
                          O.cell_keys(np.ascontiguousarray(pos_h[:, sel]), cfg))
    p1, v1 = pos.clone(), vel.clone()
    pm.step(p1, v1, 0.01, 0.00099)
    assert float(p1.min()) >= 0.0 and float(p1.max()) <= 512.0
    assert torch.isfinite(v1).all()
    p2, v2 = pos.clone(), vel.clone()
    pm.step(p2, v2, 0.01, 0.00099)
    assert torch.equal(p1, p2) and torch.equal(v1, v2)
    # resident path: 3 steps == 3 stateless steps (to rounding of the in-cell summation order)
    state = pm.ResidentParticles(pos, vel)
    p3, v3 = pos.clone(), vel.clone()
    for k in range(3):
        a = 0.01 + k * 0.00099
        state.step(a, 0.00099)
        pm.step(p3, v3, a, 0.00099)
    pr, vr = torch.empty_like(pos), torch.empty_like(vel)
    state.store(pr, vr)
    d = torch.remainder(pr.double() - p3.double() + 256.0, 512.0) - 256.0
    assert float(d.norm() / p3.double().norm()) <= 1e-6
    assert float((vr.double() - v3.double()).norm() / v3.double().norm()) <= 1e-5
    ids = state.order().long()
    assert torch.equal(torch.sort(ids).values, torch.arange(npart, device="cuda"))
    # uniform lattice: rho == mass everywhere, no kick
    pl, vl = O.lattice_ic(256, 512, jitter=0.0)
    pl, vl = dev(pl), dev(vl)
    rho = pm.density(pl, 8.0)
    assert bool((rho == 1.0).all())     # offset-0.5 lattice: every cell gets 8 * (m/8) / 8 = 1
    pm.step(pl, vl, 0.5, 0.001)
    assert float(vl.abs().max()) <= 1e-5


def test_full_size_parity_against_the_oracle_256_on_512(pm):
    """BASELINE configs[1] at FULL size -- the size bench.py times -- three free-running steps against
    the oracle (which itself reproduces the reference's digests at this size:
    tests/golden/c2_256_512_sha256.json), through BOTH product paths: the stateless drop-in calls
    (pm.density / pm.potential / pm.step) and the resident cell-ordered state bench.py measures
    (ResidentParticles -> pm_step_resident: incremental sort, k_deposit_rows, the 512-point FFT kernels,
    k_gather_tiled<512>).  Tolerances of north_star: 1e-5 relative L2 on density, potential, positions
    (periodic) and velocities, per step.  Single-axis `pos == N_CELLS` particles (~1.5 such
    events per step occur naturally at this size) are planted on each axis so the quirk is exercised at
    full size for certain.  Needs ~13 GB of host memory and ~1 minute of host time."""
    cfg = O.Config(N_CELLS=512, N_PARTS=256, STEPS=1000, N_CPU=O.max_threads())
    cfg1 = O.Config(**{**cfg.__dict__, "N_CPU": 1})           # density in particle order: the deterministic reference result
    pm.set_config(cfg_ns(cfg))
    pos_h, vel_h = O.lattice_ic(256, 512, seed=38, jitter=2.0, vel_rms=0.05)      # the digest case's input
    npart = pos_h.shape[1]
    planted = {0: [11, 5000011], 1: [77, 9000077], 2: [123, 16000123]}        # axis -> particle indices
    for axis, idx in planted.items():
        pos_h[axis, idx] = np.float32(512.0)
    pos, vel = dev(pos_h), dev(vel_h)
    state = pm.ResidentParticles(pos, vel)
    rp, rv = torch.empty_like(pos), torch.empty_like(vel)
    fg_o = O.fourier_grid(cfg)
    fg = pm.fourier_grid()
    rho = torch.empty((512, 512, 512), dtype=torch.float32, device="cuda")
    rho_r = torch.empty_like(rho)
    da = (cfg.A_END - cfg.A_INIT) / cfg.STEPS
    a = cfg.A_INIT
    phi_o = np.empty((512, 512, 512), dtype=np.float32)

    def pos_err(t):
        d = (t.cpu().numpy().astype(np.float64) - pos_h + 256.0) % 512.0 - 256.0
        return np.linalg.norm(d) / np.linalg.norm(pos_h.astype(np.float64))

    for s in range(3):
        keys_o = O.cell_keys(pos_h, cfg)
        rho_o = O.density(pos_h, 8.0, cfg1)
        if s == 0:
            # the planted particles deposit -(Nc-1)*m*w and +Nc*m*w (Q4): the mesh holds negative cells
            assert rho_o.min() < -100.0
            keys = torch.empty(npart, dtype=torch.int32, device="cuda")
            plan = pm._runtime.get_plan(512, npart, 0)
            pm._runtime.check(pm._runtime.lib().pm_cell_keys(plan.handle, pos.data_ptr(), npart, keys.data_ptr(), None), "keys")
            assert np.array_equal(keys.cpu().numpy().astype(np.int64), keys_o)        # bit-exact at full size
            del keys
        phi = pm.potential(pm.density(pos, 8.0), fg, a).cpu().numpy()
        O.advance_time(rho_o, pos_h, vel_h, fg_o, a, da, cfg, phi_out=phi_o)
        assert rel_l2(phi - phi.mean(dtype=np.float64), phi_o.astype(np.float64) - phi_o.mean(dtype=np.float64)) <= REL_L2, \
            f"potential step {s}"
        del phi
        pm.step(pos, vel, a, da, rho_out=rho)                     # stateless path, caller order
        state.step(a, da, rho_out=rho_r)                          # resident path (what bench.py times)
        state.store(rp, rv)
        for name, r_, p_, v_ in (("stateless", rho, pos, vel), ("resident", rho_r, rp, rv)):
            assert rel_l2(r_.cpu().numpy(), rho_o) <= REL_L2, f"{name}: density step {s}"
            assert pos_err(p_) <= REL_L2, f"{name}: positions step {s}"
            assert rel_l2(v_.cpu().numpy(), vel_h) <= REL_L2, f"{name}: velocities step {s}"
        a += da
    # the resident run really took the fast path: incremental sort after the first step
    assert state.sort_stats()[2] == "incremental"
    del state
    pm.release_plans()
    torch.cuda.empty_cache()


def test_tiled_gather_512_equals_flat_gather_bit_for_bit(pm):
    """k_gather_tiled<512,...> -- the kernel bench.py's roofline line is about -- against
    k_gather_kick_drift at the headline size, 256^3 particles on 512^3 cells: positions, velocities and
    storage order bit for bit over three resident steps, with a blob that overflows the staging
    capacity of a (z, row-block) step and particles sitting exactly on z == N_CELLS."""
    cfg = O.Config(N_CELLS=512, N_PARTS=256, STEPS=1000)
    pm.set_config(cfg_ns(cfg))
    rt = pm._runtime
    g = torch.Generator().manual_seed(7)
    pos_c = torch.rand((3, 256 ** 3), generator=g) * 512.0
    vel_c = torch.randn((3, 256 ** 3), generator=g) * 0.3
    pos_c[:, :40000] = torch.remainder(256.0 + 0.7 * torch.randn((3, 40000), generator=g), 512.0)
    pos_c[2, 40000:40100] = 512.0
    pos_c[0, 40100:40200] = 512.0
    out = {}
    for tiled in (0, 1):
        pm.release_plans()
        p, v = pos_c.cuda(), vel_c.cuda()
        st = pm.ResidentParticles(p, v)
        rt.check(rt.lib().pm_plan_set_gather_tiled(st.plan.handle, tiled), "tiled")
        a, da = 0.02, 0.00099
        for _ in range(3):
            st.step(a, da)
            a += da
        if tiled:
            assert st.block_stats()["blocks_over_cap"] > 0       # the overflow path really ran
        st.store(p, v)
        out[tiled] = (p.cpu(), v.cpu(), st.order().cpu())
        del st, p, v
    pm.release_plans()
    torch.cuda.empty_cache()
    assert torch.equal(out[0][0], out[1][0])
    assert torch.equal(out[0][1], out[1][1])
    assert torch.equal(out[0][2], out[1][2])


def test_full_run_power_spectrum_64_on_128(pm):
    """BASELINE config 1 (64^3 on 128^3, STEPS=100 -> 99 iterations): free-running CUDA path vs
    the oracle from identical initial conditions; P(k) of the final density within 0.1 %."""
    cfg = O.Config(N_CELLS=128, N_PARTS=64, STEPS=100, N_CPU=O.max_threads())
    pm.set_config(cfg_ns(cfg))
    pos_h, vel_h = O.lattice_ic(64, 128, seed=38, vel_rms=0.02)
    pos, vel = dev(pos_h), dev(vel_h)
    fg_o = O.fourier_grid(cfg)
    cfg1 = O.Config(**{**cfg.__dict__, "N_CPU": 1})
    nsteps = 0
    for a, da in pm.loop_scale_factors(cfg_ns(cfg)):
        pm.step(pos, vel, a, da)
        rho_o = O.density(pos_h, 8.0, cfg1)
        O.advance_time(rho_o, pos_h, vel_h, fg_o, a, da, cfg)
        nsteps += 1
        if nsteps == 10:   # the 10-step horizon of the north_star
            assert rel_l2_periodic(pos.cpu().numpy(), pos_h, 128) <= REL_L2
            assert rel_l2(vel.cpu().numpy(), vel_h) <= REL_L2
    assert nsteps == 99
    # late-time (clustered, a ~ 1) accelerations and potential, teacher-forced on the oracle state
    import importlib
    mod = importlib.import_module("cosmological_particle_mesh_simulation_b200.integrate")
    a, da = 0.99, 0.0099
    f_a1 = float(O.f(a + da, [cfg.H0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0]))
    rho_o = O.density(pos_h, 8.0, cfg1)
    phi_o = O.potential(rho_o, fg_o, a, cfg)
    acc_o = np.zeros((3, pos_h.shape[1]))
    O.integrate(pos_h.copy(), vel_h.copy(), a, f_a1, da, phi_o, cfg, acc=acc_o)
    pd = dev(pos_h)
    phi_g = pm.potential(pm.density(pd, 8.0), pm.fourier_grid(), a)
    acc_g = torch.zeros((3, pos_h.shape[1]), dtype=torch.float32, device="cuda")
    mod._integrate_device(pd, dev(vel_h), a, f_a1, da, phi_g, acc=acc_g)
    assert rel_l2(phi_g.cpu().numpy(), phi_o) <= REL_L2
    assert rel_l2(acc_g.cpu().numpy(), acc_o) <= REL_L2
    rho_gpu = pm.density(pos, 8.0).cpu().numpy()
    rho_cpu = O.density(pos_h, 8.0, cfg1)
    _, p_gpu = O.power_spectrum(rho_gpu)
    _, p_cpu = O.power_spectrum(rho_cpu)
    assert np.max(np.abs(p_gpu / p_cpu - 1.0)) <= 1e-3


@pytest.mark.parametrize("n_cells,n_parts,nblob", [(128, 64, 6000), (256, 96, 6000), (128, 64, 120000), (256, 96, 400000)])
def test_tiled_gather_equals_flat_gather_bit_for_bit(pm, n_cells, n_parts, nblob):
    """pm_gather_tiled.cuh (phi staged through shared-memory slabs, particle inputs through cp.async
    rings) does the arithmetic of k_gather_kick_drift: positions, velocities and the mover counts
    the incremental sort consumes must agree bit for bit over several resident steps, including a
    clustered blob that overflows the per-step staging capacity."""
    cfg = O.Config(N_CELLS=n_cells, N_PARTS=n_parts, STEPS=100)
    pm.set_config(cfg_ns(cfg))
    rt = pm._runtime
    pos_h, vel_h = O.lattice_ic(n_parts, n_cells, seed=5, vel_rms=0.3)
    rs = np.random.RandomState(3)
    blob = rs.normal(n_cells / 2, 0.7, size=(3, nblob)).astype(np.float32) % n_cells
    pos_h[:, :nblob] = blob                     # > CAP particles in a few (z, y-block)s; the large blobs put tens of
    #                                             thousands in ONE plane of a row block: heavy work items, cut into ranges
    pos_h[2, nblob:nblob + 100] = np.float32(n_cells)   # Q4: z == N_CELLS files under plane 0
    out = {}
    for tiled in (0, 1):
        pm.release_plans()
        p, v = dev(pos_h.copy()), dev(vel_h.copy())
        st = pm.ResidentParticles(p, v)
        rt.check(rt.lib().pm_plan_set_gather_tiled(st.plan.handle, tiled), "tiled")
        a, da = 0.02, 0.0099
        for k in range(4):
            st.step(a, da)
            a += da
            if tiled and k == 0:                    # (the blob flies apart within a few of these steps)
                heavy, light, overflow = st.gather_items()
                assert overflow == 0 and heavy + light > 0
                if nblob > 100000:
                    assert heavy > 8, (heavy, light)    # the work list really cut the blob's columns into pieces
                # every particle in exactly one item <=> the results equal the flat kernel's (below)
        st.store(p, v)
        out[tiled] = (p.cpu().numpy(), v.cpu().numpy())
    pm.release_plans()
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
