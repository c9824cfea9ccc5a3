"""The driver loop, snapshots and restart on the GPU: `pmesh.run()` =
src/pmesh.py:18-79, `save_data.save_file/from_file` = src/save_data.py:7-50.

Checked against the oracle's restatement of the cadence (oracle.loop_cadence, pmesh.py:56-74) and of
the unit factors (oracle.snapshot_units, save_data.py:10-11), and against stepping the same initial
conditions by hand through the step API that tests/test_gpu_parity.py pins to the reference."""
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from cosmological_particle_mesh_simulation_b200 import _hdf5 as H  # noqa: E402


def _cfg(**kw):
    d = dict(N_PARTS=16, N_CELLS=32, BOX_SIZE=100, N_CPU=1, RANDOM_SEED=38, STEPS=20, N_SAVE_FILES=5,
             N_PLOTS=4, PLOT_STEPS=False, PLOT_PROJECTIONS=False, PLOT_GRF=False, SAVE_DATA=True,
             SAVE_DENSITY=True, PRINT_STATUS=False, RESTART=False, RESTART_FROM_N=0, POWER=1.0,
             LCDM_TRANSFER_FUNCTION=True, OMEGA_M0=0.31, OMEGA_B0=0.04, OMEGA_K0=0.0, OMEGA_LAMBDA0=0.69,
             H0=0.68, A_INIT=0.01, A_END=1.0)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.fixture()
def pm():
    import cosmological_particle_mesh_simulation_b200 as pm
    yield pm
    pm.save_data.wait()
    pm.set_config(None)
    pm.release_plans()


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_save_file_from_device_is_bit_identical_to_the_reference_arithmetic(pm, tmp_path, monkeypatch):
    cfg = _cfg()
    pm.set_config(cfg)
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(3)
    np3 = cfg.N_PARTS ** 3
    pos_h = (rng.random((3, np3)) * cfg.N_CELLS).astype(np.float32)
    vel_h = rng.normal(size=(3, np3)).astype(np.float32)
    rho_h = rng.random((cfg.N_CELLS,) * 3, dtype=np.float32)
    pos, vel, rho = (torch.from_numpy(x).cuda() for x in (pos_h, vel_h, rho_h))
    a = 0.2575
    pm.save_file(rho, pos, vel, 3, a)
    pos.zero_(), vel.zero_(), rho.zero_()          # the caller may reuse its buffers immediately
    pm.save_data.wait()
    hf = H.Reader("Data/data.3.hdf5")
    ucp, ucv = O.snapshot_units(a, cfg)
    for i, n in enumerate(["x1", "x2", "x3"]):
        assert np.array_equal(hf[n], pos_h[i] * ucp)
    for i, n in enumerate(["vx1", "vx2", "vx3"]):
        assert np.array_equal(hf[n], vel_h[i] * ucv)
    assert np.array_equal(hf["density"], rho_h) and float(hf["a"]) == a
    # restart read: device path == host path (== src/save_data.py:35-48, tested on CPU) bit for bit
    p_h, v_h, a_h = pm.from_file(3)
    p_d, v_d, a_d = pm.from_file(3, device=0)
    assert a_h == a_d and isinstance(a_d, np.float32)
    assert np.array_equal(p_d.cpu().numpy(), p_h) and np.array_equal(v_d.cpu().numpy(), v_h)


def test_run_follows_the_reference_cadence_and_pairs_density_with_particles(pm, tmp_path, monkeypatch, capsys):
    cfg = _cfg(PLOT_STEPS=True, PLOT_PROJECTIONS=True, PLOT_GRF=True, PRINT_STATUS=True)
    pm.set_config(cfg)
    monkeypatch.chdir(tmp_path)
    pos_f, vel_f, a_end = pm.run()
    out = capsys.readouterr().out
    saves, plots = O.loop_cadence(cfg)
    trips = O.loop_trip_count(O.Config(N_CELLS=32, N_PARTS=16, STEPS=cfg.STEPS))
    assert out.count("Save step time") == trips and "Starting the integrations..." in out
    files = sorted(f for f in os.listdir("Data") if f.endswith(".hdf5"))
    assert files == sorted(["data.0.hdf5"] + ["data.%d.hdf5" % n for _, n, _ in saves])
    for _, n in plots:
        assert os.path.exists("Data/snapshots_density%d.png" % n)
        assert os.path.exists("Data/projection_density%d.png" % n)
    assert os.path.exists("Data/snapshot_grf.png")

    # replay by hand from the IC snapshot through the stateless step API
    p0, v0, a0 = pm.from_file(0, device=0)
    assert a0 == np.float32(cfg.A_INIT)
    pos, vel = p0.clone(), v0.clone()
    rho = torch.empty((32, 32, 32), dtype=torch.float32, device="cuda")
    da = (cfg.A_END - cfg.A_INIT) / cfg.STEPS
    a_current, by_iter = cfg.A_INIT, {i: (n, a) for i, n, a in saves}
    for i in range(trips):
        pm.step(pos, vel, a_current, da, rho_out=rho)
        a_current += da
        if i in by_iter:
            n, a_snap = by_iter[i]
            assert a_snap == a_current
            hf = H.Reader("Data/data.%d.hdf5" % n)
            assert float(hf["a"]) == a_current
            ucp, ucv = O.snapshot_units(a_current, cfg)
            # PRE-step density with POST-step particles (src/pmesh.py:60-67)
            assert _rel(hf["density"], rho.cpu().numpy()) < 1e-5
            d = (hf["x1"].astype(np.float64) / ucp - pos[0].cpu().numpy() + 16) % 32 - 16
            assert np.linalg.norm(d) / np.linalg.norm(pos[0].cpu().numpy().astype(np.float64)) < 1e-5
            assert _rel(hf["vx3"] / ucv, vel[2].cpu().numpy()) < 1e-4
    assert a_end == a_current
    d = (pos_f.cpu().numpy().astype(np.float64) - pos.cpu().numpy() + 16) % 32 - 16
    assert np.linalg.norm(d) / np.linalg.norm(pos.cpu().numpy().astype(np.float64)) < 1e-5


def test_restart_continues_from_a_snapshot(pm, tmp_path, monkeypatch):
    cfg = _cfg(SAVE_DENSITY=False)
    pm.set_config(cfg)
    monkeypatch.chdir(tmp_path)
    pm.run(max_steps=9)                     # writes data.0 .. data.2
    pm.save_data.wait()
    assert os.path.exists("Data/data.2.hdf5")
    cfg.RESTART, cfg.RESTART_FROM_N, cfg.SAVE_DATA = True, 2, False
    pos_r, vel_r, a_r = pm.run(max_steps=3)
    # by hand: src/pmesh.py:39-43 then three loop bodies, a_current carried as from_file returns it
    pos, vel, a_current = pm.from_file(2, device=0)
    da = (cfg.A_END - cfg.A_INIT) / cfg.STEPS
    for _ in range(3):
        pm.step(pos, vel, a_current, da)
        a_current += da
    assert type(a_r) is type(a_current) and a_r == a_current
    d = (pos_r.cpu().numpy().astype(np.float64) - pos.cpu().numpy() + 16) % 32 - 16
    assert np.linalg.norm(d) / np.linalg.norm(pos.cpu().numpy().astype(np.float64)) < 1e-5
    assert _rel(vel_r.cpu().numpy(), vel.cpu().numpy()) < 1e-4


@pytest.mark.parametrize("P", [1, 2])
def test_slab_driver_writes_the_snapshots_of_the_single_gpu_driver(pm, tmp_path, monkeypatch, P):
    """pmesh.run_slabs (initial conditions per slab, slab steps, cadence, snapshots assembled from the
    slabs) against pmesh.run on the same configuration: same files, same contents to the tolerance
    the slab step is held to, then a restart of the slab driver from one of them."""
    cfg = _cfg(PLOT_STEPS=True)
    pm.set_config(cfg)
    one, many = tmp_path / "one", tmp_path / "many"
    one.mkdir(), many.mkdir()
    monkeypatch.chdir(one)
    _, _, a_one = pm.run()
    pm.save_data.wait()
    monkeypatch.chdir(many)
    comm = pm.slab.LocalComm(P)
    ranks, a_many = pm.run_slabs(comm, peers=(P > 1))
    pm.save_data.wait()
    assert a_many == a_one
    assert sum(r.count for r in ranks) == cfg.N_PARTS ** 3
    files = sorted(f for f in os.listdir(one / "Data"))
    assert files == sorted(f for f in os.listdir(many / "Data")) and "data.0.hdf5" in files
    for f in files:
        if not f.endswith(".hdf5"):
            continue
        h1, h2 = H.Reader(str(one / "Data" / f)), H.Reader(str(many / "Data" / f))
        assert float(h1["a"]) == float(h2["a"])
        assert _rel(h2["density"], h1["density"]) < 1e-5, f
        box = cfg.N_CELLS * O.snapshot_units(float(h1["a"]), cfg)[0]
        for n in ("x1", "x2", "x3"):
            a1, a2 = h1[n].astype(np.float64), h2[n].astype(np.float64)
            d = np.minimum(np.abs(a1 - a2), box - np.abs(a1 - a2))     # a particle may sit on either side of the wrap
            assert np.linalg.norm(d) / np.linalg.norm(a1) < 1e-4, (f, n)
        for n in ("vx1", "vx2", "vx3"):
            assert _rel(h2[n], h1[n]) < 1e-3, (f, n)
    if P > 1:
        pm.slab.release_peers(ranks, comm)
    for r in ranks:
        r.close()
    # restart of the slab driver from snapshot 2: three more steps == the single-GPU restart
    cfg.RESTART, cfg.RESTART_FROM_N, cfg.SAVE_DATA, cfg.PLOT_STEPS = True, 2, False, False
    monkeypatch.chdir(one)
    pos_r, vel_r, a_r = pm.run(max_steps=3)
    monkeypatch.chdir(many)
    comm = pm.slab.LocalComm(P)
    ranks, a_s = pm.run_slabs(comm, max_steps=3, peers=False)
    assert a_s == a_r and type(a_s) is type(a_r)
    pos_s, vel_s = pm.slab.collect(ranks, comm, cfg.N_PARTS ** 3)
    d = (pos_s.double() - pos_r.double() + 16) % 32 - 16
    assert float(d.norm() / pos_r.double().norm()) < 1e-4
    assert _rel(vel_s.cpu().numpy(), vel_r.cpu().numpy()) < 1e-3
    for r in ranks:
        r.close()
