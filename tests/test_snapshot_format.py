"""CPU-side checks of the snapshot format and the driver's host logic:

* `_hdf5` writes the HDF5 subset h5py's defaults produce (src/save_data.py:16-27) and reads it
  back; the reader is also run on a file written by the real libhdf5 (a MATLAB v7.3 file that
  ships with SciPy -- the only libhdf5 output available offline), and the structures the writer
  emits are compared field by field with the ones in that file;
* `save_file` / `from_file` with NumPy arrays reproduce the arithmetic of src/save_data.py:10-24,
  :35-48 bit for bit (restated inline below with the line cites);
* the pre-step cadence prediction of `pmesh.run` equals the reference's post-step tests
  (src/pmesh.py:63-74);
* both are also pinned to the reference's OWN code: tests/golden/save_data8.npz holds what its
  unmodified save_file()/from_file() produce (stand-in h5py module), tests/golden/cadence_*.json the
  calls its unmodified simulator() makes (oracle/make_golden.py, main_driver).
"""
import os
import struct
import types
import zlib

import numpy as np
import pytest

from cosmological_particle_mesh_simulation_b200 import _hdf5 as H
from oracle import oracle as O

SCIPY_SAMPLE = None
try:
    import scipy.io.matlab as _m
    _p = os.path.join(os.path.dirname(_m.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if os.path.exists(_p):
        SCIPY_SAMPLE = _p
except Exception:  # pragma: no cover
    pass


def _cfg(**kw):
    d = dict(N_PARTS=8, N_CELLS=16, BOX_SIZE=100, N_CPU=1, RANDOM_SEED=38, STEPS=100, N_SAVE_FILES=100,
             N_PLOTS=100, PLOT_STEPS=False, PLOT_PROJECTIONS=False, PLOT_GRF=False, SAVE_DATA=True,
             SAVE_DENSITY=False, PRINT_STATUS=False, RESTART=False, RESTART_FROM_N=0, POWER=1.0,
             LCDM_TRANSFER_FUNCTION=True, OMEGA_M0=0.31, OMEGA_B0=0.04, OMEGA_K0=0.0, OMEGA_LAMBDA0=0.69,
             H0=0.68, A_INIT=0.01, A_END=1.0)
    d.update(kw)
    return types.SimpleNamespace(**d)


# ------------------------------------------------------------------------------------------------
# container format
# ------------------------------------------------------------------------------------------------
def test_round_trip_all_supported_types(tmp_path):
    rng = np.random.default_rng(1)
    data = {
        "density": rng.random((6, 5, 4), dtype=np.float32),
        "x1": rng.random(1001, dtype=np.float32),
        "f64": rng.random((3, 7)),
        "a": np.float64(0.123456789),
        "a32": np.float32(0.5),
        "ids": np.arange(17, dtype=np.int32),
        "keys": np.arange(9, dtype=np.uint32)[::-1].copy(),
        "big": np.array([2 ** 40, -3], dtype=np.int64),
        "empty": np.zeros(0, dtype=np.float32),
        "strided": rng.random((8, 8), dtype=np.float32)[::2, 1::3],
    }
    path = str(tmp_path / "t.hdf5")
    H.write(path, data)
    r = H.Reader(path)
    assert sorted(r.keys()) == sorted(data)
    for k, v in data.items():
        got = r[k]
        assert got.dtype == np.asarray(v).dtype and got.shape == np.asarray(v).shape, k
        assert np.array_equal(got, v), k
    assert set(H.check(path)) == set(data)                         # passes the structural validator
    assert r["a"].shape == () and float(r.get("a")) == 0.123456789   # np.float32(hf.get('a')) works
    assert np.float32(r.get("a")) == np.float32(0.123456789)


def test_layout_matches_the_format_specification(tmp_path):
    path = str(tmp_path / "t.hdf5")
    names = ["x1", "x2", "x3", "vx1", "vx2", "vx3", "a", "density"]      # src/save_data.py:17-25
    H.write(path, [(n, np.full(3, i, dtype=np.float32)) for i, n in enumerate(names)])
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    ver, fsv, rgv, _, shv, so, sl, _, leaf_k, int_k, flags = struct.unpack_from("<BBBBBBBBHHI", b, 8)
    assert (ver, fsv, rgv, shv, so, sl, flags) == (0, 0, 0, 0, 8, 8, 0)
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == H.UNDEF and drv == H.UNDEF and eof == len(b)
    name_off, root_hdr, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert name_off == 0 and cache == 1
    # root object header: version 1, one symbol-table message naming the same B-tree and heap
    assert b[root_hdr] == 1 and struct.unpack_from("<H", b, root_hdr + 2)[0] == 1
    mtype, msize = struct.unpack_from("<HH", b, root_hdr + 16)
    assert (mtype, msize) == (0x11, 16) and struct.unpack_from("<QQ", b, root_hdr + 24) == (btree, heap)
    # B-tree: one leaf node, group type, one child, key 0 = empty string
    assert b[btree:btree + 4] == b"TREE" and b[btree + 4] == 0 and b[btree + 5] == 0
    used, left, right, key0, child, key1 = struct.unpack_from("<HQQQQQ", b, btree + 6)
    assert used == 1 and left == H.UNDEF and right == H.UNDEF and key0 == 0
    # heap: free list ends with 1 (libhdf5's H5HL_FREE_NULL), offset 0 is the empty name
    assert b[heap:heap + 4] == b"HEAP" and b[heap + 4] == 0
    hsize, hfree, hdata = struct.unpack_from("<QQQ", b, heap + 8)
    assert hfree < hsize and struct.unpack_from("<QQ", b, hdata + hfree) == (1, hsize - hfree)
    assert b[hdata:hdata + 8] == b"\0" * 8
    # symbol-table node: entries sorted by name (strcmp), key 1 = the largest name
    assert b[child:child + 4] == b"SNOD" and b[child + 4] == 1
    count = struct.unpack_from("<H", b, child + 6)[0]
    assert count == len(names) <= 2 * leaf_k
    got = []
    for i in range(count):
        off, hdr, ctype = struct.unpack_from("<QQI", b, child + 8 + 40 * i)
        end = b.index(b"\0", hdata + off)
        got.append(b[hdata + off:end].decode())
        assert ctype == 0 and hdr % 8 == 0 and b[hdr] == 1
    assert got == sorted(names)
    end = b.index(b"\0", hdata + key1)
    assert b[hdata + key1:end].decode() == got[-1]
    # every message in every dataset header is 8-byte aligned and sized
    r = H.Reader(path)
    for n in names:
        hdr = r.entries[n]
        nmsg, _, size = struct.unpack_from("<HII", b, hdr + 2)
        o, seen = hdr + 16, []
        for _ in range(nmsg):
            t, s = struct.unpack_from("<HH", b, o)
            assert s % 8 == 0
            seen.append(t)
            o += 8 + s
        assert o == hdr + 16 + size and seen == [0x1, 0x3, 0x5, 0x8]


def test_writer_limits(tmp_path):
    with H.Writer(str(tmp_path / "a.hdf5")) as w:
        w.create_dataset("x", np.zeros(2, np.float32))
        with pytest.raises(ValueError):
            w.create_dataset("x", np.zeros(2, np.float32))
        with pytest.raises(ValueError):
            w.create_dataset("c", np.zeros(2, np.complex64))
        with pytest.raises(ValueError):
            w.create_dataset("a/b", np.zeros(2, np.float32))
        for i in range(2 * H.LEAF_K - 1):
            w.create_dataset("d%02d" % i, np.zeros(1, np.float32))
        with pytest.raises(ValueError):
            w.create_dataset("one_too_many", np.zeros(1, np.float32))
    assert len(H.Reader(str(tmp_path / "a.hdf5")).keys()) == 2 * H.LEAF_K
    with pytest.raises(ValueError):
        (tmp_path / "junk").write_bytes(b"not hdf5" * 200)
        H.Reader(str(tmp_path / "junk"))


@pytest.mark.skipif(SCIPY_SAMPLE is None, reason="SciPy's libhdf5-written sample file is not installed")
def test_reader_on_a_file_written_by_libhdf5(tmp_path):
    r = H.Reader(SCIPY_SAMPLE)                 # 512-byte user block, v0 superblock, old-style group
    assert r.keys() == ["testdouble"] and r.base == 512
    x = r["testdouble"]
    assert x.dtype == np.float64 and np.allclose(x.ravel(), np.arange(9) * np.pi / 4)
    # the same structures, byte for byte, in a file of ours holding the same dataset name:
    path = str(tmp_path / "ours.hdf5")
    H.write(path, {"testdouble": x})
    ours, ref = open(path, "rb").read(), open(SCIPY_SAMPLE, "rb").read()[512:]
    # datatype message body (IEEE float64 little endian) and dataspace message are identical
    def message(buf, reader, mtype):
        return next(bytes(body) for t, body in reader._messages(reader.entries["testdouble"]) if t == mtype)
    mine = H.Reader(path)
    assert message(ours, mine, 0x3)[:20] == message(ref, r, 0x3)[:20]
    assert message(ours, mine, 0x1) == message(ref, r, 0x1)[:8] + struct.pack("<QQ", 9, 1)
    # heap free-list convention and B-tree header agree with libhdf5's
    heap_ref = ref.index(b"HEAP")
    fr = struct.unpack_from("<Q", ref, heap_ref + 16)[0]
    data_ref = struct.unpack_from("<Q", ref, heap_ref + 24)[0]
    assert struct.unpack_from("<Q", ref, data_ref + fr)[0] == 1
    t_ref, t_ours = ref.index(b"TREE"), ours.index(b"TREE")
    assert ref[t_ref:t_ref + 32] == ours[t_ours:t_ours + 32]       # leaf, 1 entry, no siblings, key 0
    s_ref, s_ours = ref.index(b"SNOD"), ours.index(b"SNOD")
    assert ref[s_ref:s_ref + 8] == ours[s_ours:s_ours + 8]


def test_validator_accepts_libhdf5_output_and_rejects_damaged_files(tmp_path):
    """`_hdf5.check` applies the consistency rules of libhdf5's decoders.  It must accept the file
    libhdf5 itself wrote (so it is not stricter than the library) and our files, and reject ours
    once any of the fields those rules guard is damaged."""
    if SCIPY_SAMPLE is not None:
        assert list(H.check(SCIPY_SAMPLE)) == ["testdouble"]
    path = str(tmp_path / "ok.hdf5")
    names = ["x1", "x2", "x3", "vx1", "vx2", "vx3", "a", "density"]
    H.write(path, [(n, np.full(5, i, dtype=np.float32)) for i, n in enumerate(names)])
    assert sorted(H.check(path)) == sorted(names)
    good = open(path, "rb").read()
    r = H.Reader(path)
    heap = good.index(b"HEAP")
    snod = good.index(b"SNOD")
    hdr = r.entries["x2"]

    def damaged(label, edit):
        b = bytearray(good)
        edit(b)
        q = str(tmp_path / (label + ".hdf5"))
        open(q, "wb").write(bytes(b))
        with pytest.raises(ValueError):
            H.check(q)

    damaged("msg_size", lambda b: struct.pack_into("<H", b, hdr + 16 + 2, 12))            # message size not 8-aligned
    damaged("nmsgs", lambda b: struct.pack_into("<H", b, hdr + 2, 5))                      # message count mismatch
    damaged("free_undef", lambda b: struct.pack_into("<Q", b, heap + 16, H.UNDEF))         # free list must end with 1
    damaged("snod_count", lambda b: struct.pack_into("<H", b, snod + 6, 2 * H.LEAF_K + 1)) # too many symbols
    def swap(b):                                                                           # names out of order
        e0, e1 = bytes(b[snod + 8:snod + 48]), bytes(b[snod + 48:snod + 88])
        b[snod + 8:snod + 48], b[snod + 48:snod + 88] = e1, e0
    damaged("unsorted", swap)
    lay = good.index(struct.pack("<HH", 0x8, 24), hdr)                                      # layout message of x2
    damaged("layout_size", lambda b: struct.pack_into("<Q", b, lay + 8 + 10, 24))          # 5 x float32 is 20 bytes
    damaged("layout_addr", lambda b: struct.pack_into("<Q", b, lay + 8 + 2, len(good) - 8))  # data past the end
    damaged("overlap", lambda b: struct.pack_into("<Q", b, lay + 8 + 2, hdr))              # data on top of a header
    damaged("truncated", lambda b: b.__delitem__(slice(len(b) - 16, len(b))))              # shorter than the EOF address
    damaged("float_bias", lambda b: struct.pack_into("<I", b, good.index(bytes([0x11, 0x20, 31, 0]), hdr) + 16, 126))


# ------------------------------------------------------------------------------------------------
# save_file / from_file on host arrays
# ------------------------------------------------------------------------------------------------
def test_save_file_and_from_file_follow_the_reference_arithmetic(tmp_path, monkeypatch):
    import cosmological_particle_mesh_simulation_b200 as pm
    from cosmological_particle_mesh_simulation_b200 import save_data as S
    cfg = _cfg(SAVE_DENSITY=True)
    pm.set_config(cfg)
    monkeypatch.chdir(tmp_path)
    try:
        rng = np.random.default_rng(7)
        np3 = cfg.N_PARTS ** 3
        pos = (rng.random((3, np3)) * cfg.N_CELLS).astype(np.float32)
        vel = rng.normal(size=(3, np3)).astype(np.float32)
        rho = rng.random((cfg.N_CELLS,) * 3, dtype=np.float32)
        a = 0.01 + 37 * 0.0099
        S.save_file(rho, pos, vel, 5, a)
        assert os.path.exists("Data/data.5.hdf5") and not os.path.exists("Data/data.5.hdf5.part")
        hf = H.Reader("Data/data.5.hdf5")
        assert len(H.check("Data/data.5.hdf5")) == 8                 # structurally valid, 8 datasets
        unit_conv_pos, unit_conv_vel = O.snapshot_units(a, cfg)      # src/save_data.py:10-11
        assert sorted(hf.keys()) == sorted(["density", "x1", "x2", "x3", "vx1", "vx2", "vx3", "a"])
        for i, n in enumerate(["x1", "x2", "x3"]):      # :19-21
            want = pos[i] * unit_conv_pos
            assert hf[n].dtype == np.float32 and np.array_equal(hf[n], want)
        for i, n in enumerate(["vx1", "vx2", "vx3"]):   # :22-24
            assert np.array_equal(hf[n], vel[i] * unit_conv_vel)
        assert np.array_equal(hf["density"], rho) and hf["density"].shape == rho.shape
        assert hf["a"].dtype == np.float64 and hf["a"].shape == () and float(hf["a"]) == a
        # from_file, src/save_data.py:29-50 restated
        p2, v2, a2 = S.from_file(5)
        a_ref = np.float32(hf.get("a"))
        ucv = O.snapshot_units(a_ref, cfg)[1]
        assert isinstance(a2, np.float32) and a2 == a_ref
        for i, n in enumerate(["x1", "x2", "x3"]):
            assert np.array_equal(p2[i], np.array(hf.get(n)) / unit_conv_pos)
        for i, n in enumerate(["vx1", "vx2", "vx3"]):
            assert np.array_equal(v2[i], np.array(hf.get(n)) / ucv)
        assert p2.dtype == np.float32 and p2.shape == (3, np3)
        assert np.allclose(p2, pos, rtol=3e-7) and np.allclose(v2, vel, rtol=3e-6, atol=1e-7)
        # SAVE_DENSITY off: no density dataset, rho may be None
        cfg.SAVE_DENSITY = False
        S.save_file(None, pos, vel, 6, a)
        assert "density" not in H.Reader("Data/data.6.hdf5")
        # wrong particle count is refused on restart
        cfg.N_PARTS = 4
        with pytest.raises(ValueError):
            S.from_file(5)
    finally:
        pm.set_config(None)


# ------------------------------------------------------------------------------------------------
# images
# ------------------------------------------------------------------------------------------------
def test_png_encoder_and_colour_map(tmp_path):
    import torch
    from cosmological_particle_mesh_simulation_b200 import plot_helper as P
    t = torch.tensor([[0.0, 0.2, 0.4], [0.6, 1.0, float("nan")]])
    rgb = P.colour_map(t, P._PROJECTION)
    assert rgb.dtype == torch.uint8 and tuple(rgb.shape) == (2, 3, 3)
    assert rgb[0, 0].tolist() == [0, 0, 0] and rgb[0, 1].tolist() == [70, 130, 180]       # black, steelblue
    assert rgb[0, 2].tolist() == [255, 255, 255] and rgb[1, 1].tolist() == [139, 0, 0]    # white, darkred
    assert rgb[1, 2].tolist() == [0, 0, 0]                                                # bad -> black
    path = str(tmp_path / "x.png")
    P._png(path, rgb.numpy())
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n" and struct.unpack_from(">II", b, 16) == (3, 2)
    i = b.index(b"IDAT")
    n = struct.unpack_from(">I", b, i - 4)[0]
    raw = zlib.decompress(b[i + 4:i + 4 + n])
    rows = np.frombuffer(raw, np.uint8).reshape(2, 1 + 9)
    assert np.array_equal(rows[:, 1:].reshape(2, 3, 3), rgb.numpy()) and not rows[:, 0].any()


# ------------------------------------------------------------------------------------------------
# cadence
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("steps,n_save,n_plot", [(100, 100, 100), (1000, 100, 100), (20, 5, 4), (37, 10, 3)])
def test_reference_cadence_restatement_is_consistent(steps, n_save, n_plot):
    """The driver decides BEFORE a step whether its density is needed, from a_current + da; that is
    the same floating-point value the reference tests after `a_current += da` (src/pmesh.py:63-65)."""
    cfg = _cfg(STEPS=steps, N_SAVE_FILES=n_save, N_PLOTS=n_plot)
    saves, plots = O.loop_cadence(cfg)
    da = (cfg.A_END - cfg.A_INIT) / cfg.STEPS
    da_save = (cfg.A_END - cfg.A_INIT) / cfg.N_SAVE_FILES
    a, n_file, pre = cfg.A_INIT, 1, []
    i = 0
    while a < cfg.A_END - da:
        a_next = a + da
        if a_next >= cfg.A_INIT + n_file * da_save:
            pre.append((i, n_file, a_next))
            n_file += 1
        a += da
        assert a == a_next
        i += 1
    assert pre == saves and len(saves) >= 1


# ------------------------------------------------------------------------------------------------
# property test: any set of supported arrays survives the container
# ------------------------------------------------------------------------------------------------
try:
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp
    HAVE_HYPOTHESIS = True
except Exception:  # pragma: no cover
    HAVE_HYPOTHESIS = False

if HAVE_HYPOTHESIS:
    _dtypes = st.sampled_from([np.float32, np.float64, np.int32, np.int64, np.uint32, np.uint8, np.int16])
    _arrays = _dtypes.flatmap(lambda dt: hnp.arrays(dt, hnp.array_shapes(min_dims=0, max_dims=3, min_side=0, max_side=7)))
    _names = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCXYZ0123456789_.", min_size=1, max_size=20)

    @settings(max_examples=40, deadline=None)
    @given(st.dictionaries(_names, _arrays, min_size=0, max_size=12))
    def test_any_flat_group_of_arrays_round_trips(tmp_path_factory, data):
        path = str(tmp_path_factory.mktemp("h5") / "p.hdf5")
        H.write(path, data)
        r = H.Reader(path)
        assert sorted(r.keys()) == sorted(data)
        assert os.path.getsize(path) == struct.unpack_from("<Q", open(path, "rb").read(48), 40)[0]   # EOF address
        assert set(H.check(path)) == set(data)
        for k, v in data.items():
            got = r[k]
            assert got.dtype == v.dtype and got.shape == v.shape
            assert got.tobytes() == np.ascontiguousarray(v).tobytes()      # bit-exact, NaN payloads included


# ------------------------------------------------------------------------------------------------
# pinned to the reference's own save_data.py / pmesh.py (tests/golden/save_data8.npz, cadence_*.json,
# produced by oracle/make_golden.py main_driver() with stand-in h5py / matplotlib modules)
# ------------------------------------------------------------------------------------------------
def test_save_file_and_from_file_equal_the_reference_bit_for_bit(tmp_path, monkeypatch, golden_dir):
    import cosmological_particle_mesh_simulation_b200 as pm
    from cosmological_particle_mesh_simulation_b200 import save_data as S
    g = np.load(os.path.join(golden_dir, "save_data8.npz"))
    cfg = _cfg(N_PARTS=8, N_CELLS=16, SAVE_DENSITY=True)
    pm.set_config(cfg)
    monkeypatch.chdir(tmp_path)
    try:
        pos, vel, rho = g["pos"], g["vel"], g["rho"]
        for i, a in enumerate(g["a_values"].tolist()):
            S.save_file(rho, pos, vel, i, a)
            hf = H.Reader("Data/data.%d.hdf5" % i)
            names = sorted(k[len("file%d_" % i):] for k in g.files if k.startswith("file%d_" % i))
            assert sorted(hf.keys()) == names
            for n in names:
                want = g["file%d_%s" % (i, n)]
                got = hf[n]
                assert got.dtype == want.dtype and got.shape == want.shape, n
                assert got.tobytes() == want.tobytes(), f"dataset {n} of snapshot {i}"
            p2, v2, a2 = S.from_file(i)
            assert str(np.asarray(a2).dtype) == str(g["from%d_a_dtype" % i]) and np.asarray(a2).tobytes() == g["from%d_a" % i].tobytes()
            assert p2.tobytes() == g["from%d_pos" % i].tobytes() and v2.tobytes() == g["from%d_vel" % i].tobytes()
    finally:
        pm.set_config(None)


@pytest.mark.parametrize("name", ["cadence_100_100", "cadence_1000_100", "cadence_20_5", "cadence_37_10"])
def test_cadence_restatement_equals_the_reference_driver(golden_dir, name):
    """oracle.loop_cadence (which tests/test_gpu_driver.py holds pmesh.run() to) against the calls the
    reference's own simulator() made: which iterations saved under which index at which a_current,
    which plotted, and one status line per iteration when PRINT_STATUS."""
    import json
    d = json.load(open(os.path.join(golden_dir, name + ".json")))
    case, calls = d["case"], d["calls"]
    cfg = _cfg(STEPS=case["STEPS"], N_SAVE_FILES=case["N_SAVE_FILES"], N_PLOTS=case["N_PLOTS"], A_INIT=case["A_INIT"])
    saves, plots = O.loop_cadence(cfg)
    ref_saves = [(n, a) for kind, n, *rest in calls if kind == "save" for a in rest]
    assert ref_saves[0] == (0, cfg.A_INIT)                                   # the IC snapshot, pmesh.py:46-48
    assert [(n, a) for _, n, a in saves] == ref_saves[1:]
    if case.get("PLOT_STEPS"):
        assert [float(n) for _, n in plots] == [c[1] for c in calls if c[0] == "plot"]
    if case.get("PLOT_PROJECTIONS"):
        assert [c[1:] for c in calls if c[0] == "proj"] == [[float(n), 15.0] for _, n in plots]
    if case.get("PRINT_STATUS"):
        trips = O.loop_trip_count(O.Config(STEPS=case["STEPS"], A_INIT=case["A_INIT"]))
        assert sum(1 for c in calls if c[0] == "status") == trips


def test_projection_equals_the_reference_bitwise(golden_dir):
    """analysis.project (what plot_projection images) against the reference's own numba `project`
    (src/plot_helper.py:65-72, tests/golden/project16.npz): float64 sums of the first n planes."""
    import torch
    from cosmological_particle_mesh_simulation_b200 import analysis
    g = np.load(os.path.join(golden_dir, "project16.npz"))
    rho = torch.from_numpy(g["rho"])
    for k in [k for k in g.files if k.startswith("proj_")]:
        n = int(k.split("_")[1])
        got = analysis.project(rho, n).numpy()
        assert got.dtype == g[k].dtype == np.float64 and got.tobytes() == g[k].tobytes(), k
