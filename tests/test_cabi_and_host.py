"""CPU-side checks (no GPU needed): the C-ABI library loads and exports every symbol that
include/pmstep.h declares, fails loudly without a device, and the host-side mirror of the
reference interface (configure_me names, f(), loop predicate, flat-module drop-in) is right."""
import ctypes
import os
import re
import subprocess
import sys
import types

import numpy as np
import pytest

from oracle import oracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "cosmological_particle_mesh_simulation_b200")
HEADER = os.path.join(REPO, "include", "pmstep.h")


@pytest.fixture(scope="module")
def libpath():
    so = os.path.join(PKG, "libpmstep.so")
    if not os.path.exists(so):
        sys.path.insert(0, REPO)
        import __graft_entry__ as g
        g.build()
    assert os.path.exists(so)
    return so


def declared_symbols():
    text = open(HEADER).read()
    return re.findall(r"PM_API\s+[\w\s\*]+?\b(pm_\w+)\s*\(", text)


def test_header_declares_the_documented_surface():
    syms = declared_symbols()
    for must in ["pm_plan_create", "pm_plan_destroy", "pm_cell_keys", "pm_sort_by_cell", "pm_deposit_cic",
                 "pm_poisson", "pm_gather_kick_drift", "pm_step", "pm_step_host", "pm_fourier_grid"]:
        assert must in syms
    assert len(syms) == len(set(syms))


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in pmstep.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared_symbols()) == {s for s in exported if s.startswith("pm_")}
    # python binding table and header agree
    from cosmological_particle_mesh_simulation_b200 import _runtime as rt
    assert set(rt.EXPORTED_SYMBOLS) == set(declared_symbols())


def test_built_for_sm_100a_only(libpath):
    out = subprocess.run(["cuobjdump", "--list-elf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import cosmological_particle_mesh_simulation_b200 as pm
    lib = pm._runtime.lib()
    assert b"pmstep" in lib.pm_version()
    h = ctypes.c_void_p()
    assert lib.pm_plan_create(ctypes.byref(h), 32, 1000, -1) == -5      # PM_ERR_NO_DEVICE
    assert lib.pm_plan_create(ctypes.byref(h), 2000, 1000, -1) == -2    # PM_ERR_UNSUPPORTED
    assert lib.pm_step(None, None, None, 0, 1.0, 0.1, 0.1, 1.0, 0.3, None, None) == -1  # PM_ERR_INVALID
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pm.density(np.zeros((3, 8), np.float32), 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pm.fourier_grid()


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, fn)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), f"{fn} mentions the oracle"
    for fn in ["bench.py"]:
        pass  # bench.py may use the oracle only in its cpu_baseline / --impl reference legs


def test_configure_me_names_and_defaults_match_reference():
    from cosmological_particle_mesh_simulation_b200 import configure_me as cm
    want = dict(N_PARTS=256, N_CELLS=512, BOX_SIZE=100, N_CPU=16, RANDOM_SEED=38, STEPS=1000,
                N_SAVE_FILES=100, N_PLOTS=100, PLOT_STEPS=False, PLOT_PROJECTIONS=False, PLOT_GRF=False,
                SAVE_DATA=True, SAVE_DENSITY=False, PRINT_STATUS=True, RESTART=False, RESTART_FROM_N=0,
                POWER=1.00, LCDM_TRANSFER_FUNCTION=True, OMEGA_M0=0.31, OMEGA_B0=0.04, OMEGA_K0=0.00,
                OMEGA_LAMBDA0=0.69, H0=0.68, A_INIT=0.01, A_END=1.00)   # configure_me.py:7-40
    for k, v in want.items():
        assert getattr(cm, k) == v, k


def test_f_matches_oracle_bitwise_and_loop_predicate():
    import cosmological_particle_mesh_simulation_b200 as pm
    for a in np.linspace(0.01, 1.0, 57):
        assert pm.f(float(a), [0.68, 0.69, 0.0]) == float(O.f(float(a), [0.68, 0.69, 0.0]))
    for steps, want in [(10, 10), (100, 99), (1000, 999), (500, 500), (2000, 1999)]:
        cfg = types.SimpleNamespace(A_INIT=0.01, A_END=1.00, STEPS=steps)
        sched = pm.loop_scale_factors(cfg)
        assert len(sched) == want == O.loop_trip_count(O.Config(STEPS=steps))
        assert sched[0] == (0.01, (1.00 - 0.01) / steps)


def test_cosmology_keeps_float32_on_the_restart_path():
    """save_data.from_file returns `a` as np.float32; the reference's f(a+da) (np.sqrt, src/cosmology.py:27)
    then stays in float32 under NumPy 2.  Same here: dtype kept, bits equal to the float32 evaluation.
    Also the names `from cosmology import *` gives the reference's modules (cosmo, np)."""
    import numpy as np
    from cosmological_particle_mesh_simulation_b200 import cosmology as C
    a = np.float32(0.3) + np.float32(0.00099)
    cosm = [0.68, 0.69, 0.0]
    got = C.f(a, cosm)
    assert isinstance(got, np.floating) and got.dtype == np.float32
    want = 1 / np.sqrt((cosm[0] + cosm[2] * a + cosm[1] * a ** 3) / a)
    assert want.dtype == np.float32 and got.tobytes() == want.tobytes()
    assert float(got) != 1 / np.sqrt((cosm[0] + cosm[2] * float(a) + cosm[1] * float(a) ** 3) / float(a))  # float64 differs
    assert C.H(a, 0.68, [0.31, 0.69, 0.0]).dtype == np.float32
    # float64 / Python float input: float64 as before
    assert np.asarray(C.f(0.3, cosm)).dtype == np.float64
    c = C.cosmo()
    assert c.f0 == C.f(0.01, [0.31, 0.69, 0.0]) and c.Dt == C.Dt(0.01, [0.31, 0.69, 0.0]) and C.np is np


def test_cosmology_functions_equal_the_reference_bitwise(golden_dir):
    """tests/golden/cosmology.json: f (with the loop's argument order, SURVEY Q1, and the intended one),
    H and Dt evaluated by the reference's own cosmology.py (oracle/make_golden.py, main_driver)."""
    import json
    from cosmological_particle_mesh_simulation_b200 import cosmology as C
    g = json.load(open(os.path.join(golden_dir, "cosmology.json")))
    H0, om, ol, ok = g["H0"], g["OMEGA_M0"], g["OMEGA_LAMBDA0"], g["OMEGA_K0"]
    for r in g["rows"]:
        a = r["a"]
        assert float(C.f(a, [H0, ol, ok])) == r["f_loop"] == float(O.f(a, [H0, ol, ok]))
        assert float(C.f(a, [om, ol, ok])) == r["f"]
        assert float(C.H(a, H0, [om, ol, ok])) == r["H"]
        assert float(C.Dt(a, [om, ol, ok])) == r["Dt"]


def test_config_lookup_order():
    import cosmological_particle_mesh_simulation_b200 as pm
    pm.set_config(types.SimpleNamespace(N_CELLS=48))
    assert pm.config().N_CELLS == 48
    pm.set_config(None)
    assert pm.config().N_CELLS == 512 or "configure_me" in sys.modules


def test_flat_module_drop_in_layout():
    """The reference imports `from density import density` etc. from a flat directory
    (pmesh.py:5-16); putting the package directory on sys.path must satisfy those imports."""
    code = ("import sys; sys.path.insert(0, %r);"
            "from density import density; from integrate import advance_time, integrate;"
            "from fourier_utils import fourier_grid; from potential import potential;"
            "from cosmology import f, H, Dt; import configure_me;"
            "from zeldovich import zeldovich; from gaussian_random_field import gaussian_random_field;"   # pmesh.py:7,9
            "from save_data import save_file, from_file;"                                                 # pmesh.py:11
            "from plot_helper import plot_step, plot_grf, plot_projection;"                               # pmesh.py:12
            "import pmesh;"
            "print(configure_me.N_CELLS, density.__name__, advance_time.__name__, pmesh.simulator.__name__)") % PKG
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["512", "density", "advance_time", "simulator"]
    # `python pmesh.py` (pmesh.py:86-93) from that directory: without a GPU it must fail loudly, not fall back
    out = subprocess.run([sys.executable, "pmesh.py"], capture_output=True, text=True, cwd=PKG,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode != 0 and "no CPU fallback" in out.stderr


@pytest.mark.skipif(not os.path.exists("/root/reference/src/pmesh.py"), reason="needs the reference checkout (build container only)")
def test_unmodified_reference_driver_resolves_every_import_to_the_package(tmp_path):
    """the reference's own pmesh.py, run with the package directory ahead of
    its src/ on sys.path.  Its imports (density, integrate, zeldovich, fourier_utils,
    gaussian_random_field, cosmology, save_data, plot_helper, configure_me) must all bind to the GPU
    modules -- the reference's versions would die on `import pyfftw` / `h5py` / `matplotlib` -- and without
    a GPU the run must stop at the first device call, loudly."""
    (tmp_path / "configure_me.py").write_text(
        "N_PARTS=16; N_CELLS=32; BOX_SIZE=100; N_CPU=1; RANDOM_SEED=38; STEPS=10; N_SAVE_FILES=5; N_PLOTS=5\n"
        "PLOT_STEPS=False; PLOT_PROJECTIONS=False; PLOT_GRF=False; SAVE_DATA=True; SAVE_DENSITY=False; PRINT_STATUS=True\n"
        "RESTART=False; RESTART_FROM_N=0; POWER=1.0; LCDM_TRANSFER_FUNCTION=True\n"
        "OMEGA_M0=0.31; OMEGA_B0=0.04; OMEGA_K0=0.0; OMEGA_LAMBDA0=0.69; H0=0.68; A_INIT=0.01; A_END=1.0\n")
    code = ("import sys, runpy; sys.path[:0] = [%r, %r, '/root/reference/src'];"
            "runpy.run_path('/root/reference/src/pmesh.py', run_name='__main__')") % (str(tmp_path), PKG)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path),
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode != 0
    assert "ModuleNotFoundError" not in out.stderr, out.stderr[-1500:]
    assert "no CPU fallback" in out.stderr and os.path.join(PKG, "gaussian_random_field.py") in out.stderr
    assert "Starting the simulation for 16^3 particles with 32^3 grid cells" in out.stdout     # pmesh.py:20, its own banner


def test_lazy_drop_in_handle_mechanics(monkeypatch):
    """_session.ResidentView (what advance_time returns with set_resident_dropin("lazy")): metadata access
    leaves the deferred write-back pending, any access to data triggers it first, and the handle shares
    storage and version counter with the caller's tensor."""
    import numpy as np
    import torch
    from cosmological_particle_mesh_simulation_b200 import _session as S
    calls = []
    monkeypatch.setattr(S, "sync", lambda: calls.append(1))
    t = torch.arange(6.0).reshape(3, 2)
    v = t.as_subclass(S.ResidentView)
    v.__dict__["_pm_base"] = t
    assert S.unwrap(v) is t and S.unwrap(t) is t
    assert (tuple(v.shape), v.dtype, v.device.type, v.dim(), v.size(1), len(v), v.is_contiguous(), v.numel()) == \
        ((3, 2), torch.float32, "cpu", 2, 2, 3, True, 6)
    assert calls == []                                   # nothing above read data
    assert type(v + 1) is torch.Tensor and len(calls) == 1
    assert np.asarray(v).shape == (3, 2) and len(calls) == 2
    assert v.data_ptr() == t.data_ptr() and len(calls) == 3
    v[0, 0] = 5.0                                        # in-place write through the handle: sync first, then one version bump
    assert len(calls) == 4 and float(t[0, 0]) == 5.0 and t._version == 1


def test_raw_writes_of_the_package_are_made_visible_to_torch():
    """_session.after_raw_write: what the package's C-ABI wrappers call after writing into a caller's tensor
    through data_ptr() -- the version counter moves (views share it), None and handles are accepted;
    before_raw_access without a session is a no-op."""
    import torch
    from cosmological_particle_mesh_simulation_b200 import _session as S
    t = torch.zeros(3, 4)
    row = t[1]
    v0 = t._version
    S.after_raw_write(t, None)
    assert t._version > v0 and row._version == t._version
    v1 = t._version
    h = t.as_subclass(S.ResidentView)
    h.__dict__["_pm_base"] = t
    S.after_raw_write(h)
    assert t._version > v1
    assert S._session is None
    S.before_raw_access(t, None, h)          # nothing to flush


def test_pooled_host_arrays_stay_busy_while_any_view_lives():
    """_runtime._PinnedHolder: what NumPy sees as the owner of a pooled result buffer (density() on NumPy input
    returns such arrays) -- it must stay alive, i.e. the buffer must not be handed out again, as long as the array
    OR ANY VIEW of it exists; small arrays are never page-locked."""
    import gc
    import weakref
    import numpy as np
    import torch
    from cosmological_particle_mesh_simulation_b200 import _runtime as rt
    t = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    h = rt._PinnedHolder(t)
    alive = weakref.ref(h)
    a = np.asarray(h)
    assert a.shape == (3, 4) and a.dtype == np.float32 and a[2, 3] == 11.0
    a[0, 0] = 7.0
    assert float(t[0, 0]) == 7.0                      # the array IS the buffer, not a copy
    del h
    row = a[1]
    flat = a.reshape(-1)[2:5]
    del a
    gc.collect()
    assert alive() is not None                        # two views left
    del row
    gc.collect()
    assert alive() is not None
    del flat
    gc.collect()
    assert alive() is None                            # now the pool may reuse the buffer
    assert rt.pin_host_array(np.zeros(16, dtype=np.float32)) is False          # too small: no library call
    assert rt.pin_host_array(np.zeros((1 << 19, 2), dtype=np.float32)[:, 0]) is False   # not contiguous


def test_host_step_ranges_cover_every_particle_once():
    """pm_step_host_range: the ranges pm_step_host uploads velocities in, pushes and downloads (host arithmetic of
    the library itself, no device): consecutive, disjoint, together [0, np), and every boundary except the last a
    multiple of 64 particles (256-byte aligned rows for the 2-D copies)."""
    import ctypes
    from cosmological_particle_mesh_simulation_b200 import _runtime as rt
    L = rt.lib()
    for np_ in [0, 1, 63, 64, 65, 100, 255, 256, 257, 4096, 125000, 262144, 884736, 16777216, 16777215, (1 << 31) + 12345]:
        i0, i1, n = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int(0)
        assert L.pm_step_host_range(np_, 0, ctypes.byref(i0), ctypes.byref(i1), ctypes.byref(n)) == 0
        assert n.value >= 1
        end = 0
        for k in range(n.value):
            assert L.pm_step_host_range(np_, k, ctypes.byref(i0), ctypes.byref(i1), None) == 0
            assert i0.value == end and i1.value >= i0.value
            if k + 1 < n.value:
                assert i1.value % 64 == 0
            end = i1.value
        assert end == np_
        assert L.pm_step_host_range(np_, n.value, ctypes.byref(i0), ctypes.byref(i1), None) != 0
    assert L.pm_step_host_range(-1, 0, ctypes.byref(i0), ctypes.byref(i1), None) != 0
