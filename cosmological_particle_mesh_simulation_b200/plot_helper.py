"""Density images (reference: src/plot_helper.py:12-72).

Same function names, arguments and output file names as the reference.  The reductions and the
colour mapping run on the GPU; matplotlib is not available in this image, so the files are plain
mesh-resolution PNGs (one pixel per cell, no axes or colour bar) written by a ~20-line encoder:

    plot_step(rho, n)            Data/snapshots_density{n}.png    rho[0] linear in [0, 3*mass], viridis
    plot_grf(rho)                Data/snapshot_grf.png            rho[0] linear in [-1, max], viridis
    plot_projection(rho, n, d)   Data/projection_density{n}.png   mean of the first BOX_SIZE/d planes,
                                                                  log scale in [0.2, 25*mass], the
                                                                  reference's six-colour map
    project(rho, n_slices)       float64 sum of the first n_slices planes (plot_helper.py:65-72)
"""
import os
import struct
import zlib

import numpy as np
import torch

try:
    from . import _runtime as rt
    from .analysis import project
except ImportError:  # flat layout
    import _runtime as rt
    from analysis import project

DATA_DIR = "Data/"

# anchor colours, evenly spaced like matplotlib's LinearSegmentedColormap.from_list
_VIRIDIS = [(68, 1, 84), (72, 40, 120), (62, 74, 137), (49, 104, 142), (38, 130, 142), (31, 158, 137),
            (53, 183, 121), (109, 205, 89), (180, 222, 44), (253, 231, 37)]
_PROJECTION = [(0, 0, 0), (70, 130, 180), (255, 255, 255), (255, 255, 0), (255, 165, 0), (139, 0, 0)]
# names of plot_helper.py:36: black, steelblue, white, yellow, orange, darkred


def _png(path, rgb):
    """rgb: uint8 [H, W, 3] NumPy array -> 8-bit truecolour PNG."""
    h, w, _ = rgb.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))

    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def colour_map(t, anchors, bad=(0, 0, 0)):
    """t: tensor of normalised values (0..1; NaN/inf = `bad`) -> uint8 [..., 3] on t's device."""
    lut = torch.tensor(anchors, dtype=torch.float32, device=t.device)
    invalid = ~torch.isfinite(t)
    x = torch.nan_to_num(t.to(torch.float32), nan=0.0, posinf=1.0, neginf=0.0).clamp(0.0, 1.0) * (len(anchors) - 1)
    i0 = x.floor().clamp(max=len(anchors) - 2).long()
    w = (x - i0).unsqueeze(-1)
    rgb = (lut[i0] * (1 - w) + lut[i0 + 1] * w).round().to(torch.uint8)
    rgb[invalid] = torch.tensor(bad, dtype=torch.uint8, device=t.device)
    return rgb


def _as_device(rho):
    if isinstance(rho, torch.Tensor) and rho.is_cuda:
        return rho
    return rt.to_device(np.ascontiguousarray(rho, dtype=np.float32), rt.current_device())


def plot_step(rho, savestep):
    """plot_helper.py:12-21."""
    cfg = rt.config()
    mass = (cfg.N_CELLS / cfg.N_PARTS) ** 3
    img = _as_device(rho)[0] / (mass * 3)
    _png(os.path.join(DATA_DIR, "snapshots_density{}.png".format(savestep)), colour_map(img, _VIRIDIS).cpu().numpy())


def plot_grf(rho):
    """plot_helper.py:23-30."""
    r = _as_device(rho)[0]
    vmax = r.max()
    img = (r + 1.0) / (vmax + 1.0)
    _png(os.path.join(DATA_DIR, "snapshot_grf.png"), colour_map(img, _VIRIDIS).cpu().numpy())


def plot_projection(rho, savestep, depth):
    """plot_helper.py:32-63: LogNorm(vmin=0.2, vmax=25*mass) of projection/n_slices; cells at or
    below zero are 'bad' (black)."""
    cfg = rt.config()
    print("Plotting projection projection_density{}.png".format(np.int32(savestep)))
    mass = (cfg.N_CELLS / cfg.N_PARTS) ** 3
    n_slices = np.int32(cfg.BOX_SIZE / depth)
    mean = project(_as_device(rho), n_slices) / float(n_slices)
    lo, hi = np.log(0.2), np.log(mass * 25)
    img = (torch.log(mean) - lo) / (hi - lo)
    img = torch.where(mean > 0, img, torch.full_like(img, float("nan")))
    _png(os.path.join(DATA_DIR, "projection_density{}.png".format(np.int32(savestep))),
         colour_map(img, _PROJECTION).cpu().numpy())
