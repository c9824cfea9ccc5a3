"""Snapshot / restart files (reference: src/save_data.py:7-50).

Same function names, arguments, file names, dataset names, dtypes and unit factors as the
reference: ``save_file(rho, positions, velocities, step, a)`` writes ``Data/data.{step}.hdf5`` with
float32 datasets x1,x2,x3 (Mpc), vx1,vx2,vx3 (km/s), the scalar float64 ``a`` and, if
SAVE_DENSITY, the float32 mesh ``density``; ``from_file(step)`` returns ``(positions, velocities,
a)`` in code units with ``a`` a numpy.float32, exactly like save_data.py:29-50.

What is different is how the bytes get there.  h5py is not available, so the container format is
written by `_hdf5` (the subset of HDF5 that h5py's defaults produce).  For CUDA tensors the unit
conversion runs on the device, the device->host copies go to pinned memory on a copy stream, and
a background thread writes the file, so `save_file` returns as soon as the copies are queued and
the next steps overlap the transfer and the disk write (`wait()` blocks until everything is on
disk; the interpreter also waits at exit).  NumPy arrays are written synchronously, as in the
reference."""
import atexit
import os
import queue
import threading

import numpy as np
import torch

try:
    from . import _runtime as rt
    from . import _hdf5
except ImportError:  # flat layout (package directory on sys.path)
    import _runtime as rt
    import _hdf5

DATA_DIR = "Data/"     # src/save_data.py:13,16


def unit_conversions(a, cfg=None):
    """(unit_conv_pos [Mpc], unit_conv_vel [km/s]) of src/save_data.py:10-11."""
    c = cfg or rt.config()
    unit_conv_pos = 7.8 * (c.BOX_SIZE / (c.N_CELLS / 128)) / 10 ** 3
    unit_conv_vel = 0.781 * c.BOX_SIZE * c.H0 / (a * c.N_CELLS / 128)
    return unit_conv_pos, unit_conv_vel


def filename(step):
    return os.path.join(DATA_DIR, "data.{}.hdf5".format(step))


# ------------------------------------------------------------------------------------------------
# background writer
# ------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.q = queue.Queue()
        self.thread = None
        self.error = None
        self.lock = threading.Lock()

    def submit(self, path, datasets, event, keep_alive):
        with self.lock:
            if self.thread is None or not self.thread.is_alive():
                self.thread = threading.Thread(target=self._run, name="pm-snapshot-writer", daemon=True)
                self.thread.start()
        self.q.put((path, datasets, event, keep_alive))

    def _run(self):
        while True:
            job = self.q.get()
            try:
                path, datasets, event, _keep = job
                if event is not None:
                    event.synchronize()
                _write(path, [(k, v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in datasets])
            except BaseException as e:  # surfaced by wait() / the next save_file
                self.error = e
            finally:
                del job
                self.q.task_done()

    def wait(self):
        self.q.join()
        if self.error is not None:
            e, self.error = self.error, None
            raise e


_writer = _Writer()
_copy_streams = {}


def wait():
    """Block until every queued snapshot is on disk; re-raises a writer error."""
    _writer.wait()


atexit.register(lambda: _writer.q.join())


def _write(path, datasets):
    tmp = path + ".part"
    with _hdf5.Writer(tmp) as w:
        for name, arr in datasets:
            w.create_dataset(name, arr)
    os.replace(tmp, path)          # a reader never sees a half-written snapshot


def _copy_stream(dev):
    if dev not in _copy_streams:
        _copy_streams[dev] = torch.cuda.Stream(device=dev)
    return _copy_streams[dev]


# ------------------------------------------------------------------------------------------------
# save_file / from_file
# ------------------------------------------------------------------------------------------------
def save_file(rho, positions, velocities, step, a, block=None):
    """src/save_data.py:7-27.  CUDA tensors: asynchronous (see module docstring) unless
    block=True; NumPy arrays / CPU tensors: synchronous.  rho may be None when SAVE_DENSITY is off."""
    cfg = rt.config()
    print("Writing to disk: data.{}.hdf5".format(step))
    unit_conv_pos, unit_conv_vel = unit_conversions(a, cfg)
    os.makedirs(os.path.dirname(DATA_DIR), exist_ok=True)
    path = filename(step)
    save_density = bool(getattr(cfg, "SAVE_DENSITY", False))
    if save_density and rho is None:
        raise ValueError("SAVE_DENSITY is set but no density mesh was passed")
    names = [("x1", positions, 0, unit_conv_pos), ("x2", positions, 1, unit_conv_pos),
             ("x3", positions, 2, unit_conv_pos), ("vx1", velocities, 0, unit_conv_vel),
             ("vx2", velocities, 1, unit_conv_vel), ("vx3", velocities, 2, unit_conv_vel)]
    a_data = np.asarray(a)        # h5py: create_dataset('a', data=a) -> scalar dataset of a's dtype
    if a_data.dtype.kind not in "fiu":
        raise TypeError("a must be a real number")

    on_gpu = isinstance(positions, torch.Tensor) and positions.is_cuda
    if not on_gpu:
        if _writer.error is not None:
            wait()
        p, v = rt.as_host_f32(positions), rt.as_host_f32(velocities)
        datasets = []
        if save_density:
            r = rho.detach().cpu().numpy() if isinstance(rho, torch.Tensor) else np.asarray(rho)
            datasets.append(("density", r))
        for name, arr, row, conv in names:
            src = p if arr is positions else v
            datasets.append((name, src[row] * conv))          # float32 * Python float -> float32
        datasets.append(("a", a_data))
        _write(path, datasets)
        return

    if _writer.error is not None:
        wait()
    dev = positions.device.index
    main = torch.cuda.current_stream(dev)
    side = _copy_stream(dev)
    staged, keep = [], []
    with torch.cuda.device(dev):
        # conversion on the compute stream (the caller may overwrite positions right after we
        # return); the copies to pinned memory run on the copy stream behind it
        if save_density:
            keep.append(("density", rho.detach().clone()))
        for name, arr, row, conv in names:
            keep.append((name, arr[row] * conv))              # float32 tensor * Python float -> float32
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            for name, t in keep:
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t, non_blocking=True)
                t.record_stream(side)
                staged.append((name, h))
            done = torch.cuda.Event()
            done.record(side)
    staged.append(("a", a_data))
    _writer.submit(path, staged, done, keep)
    if block:
        wait()


def from_file(step, device=None):
    """src/save_data.py:29-50: (positions, velocities, a) in code units, float32 [3, N_PARTS^3]
    NumPy arrays and a numpy.float32 scale factor.  With device=<index> the two arrays are returned
    as CUDA tensors instead (divided on the device by the same float32 factors)."""
    cfg = rt.config()
    wait()
    np3 = int(cfg.N_PARTS) ** 3
    hf = _hdf5.Reader(filename(step))
    a = np.float32(hf.get("a"))
    unit_conv_pos, unit_conv_vel = unit_conversions(a, cfg)
    rows = [("x1", unit_conv_pos), ("x2", unit_conv_pos), ("x3", unit_conv_pos),
            ("vx1", unit_conv_vel), ("vx2", unit_conv_vel), ("vx3", unit_conv_vel)]
    for name, _ in rows:
        shape = hf.info(name)[0]
        if shape != (np3,):
            raise ValueError(f"{filename(step)}: dataset {name} has shape {shape}, N_PARTS^3 = {np3}")
    if device is None:
        positions = np.zeros((3, np3), dtype=np.float32)
        velocities = np.zeros((3, np3), dtype=np.float32)
        for i, (name, conv) in enumerate(rows):
            (positions if i < 3 else velocities)[i % 3] = np.array(hf.get(name)) / conv
        return positions, velocities, a
    dev = f"cuda:{int(device)}"
    out = torch.empty((2, 3, np3), dtype=torch.float32, device=dev)
    for i, (name, conv) in enumerate(rows):
        raw = torch.from_numpy(np.ascontiguousarray(hf.get(name), dtype=np.float32)).to(dev)
        # a 0-dim device divisor keeps this a true IEEE division (not a multiply by 1/conv)
        torch.div(raw, torch.tensor(np.float32(conv), device=dev), out=out[i // 3, i % 3])
    return out[0], out[1], a
