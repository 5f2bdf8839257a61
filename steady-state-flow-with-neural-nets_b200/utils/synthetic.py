"""Synthetic binary boundary images shaped like the reference's LBM training set: channel walls on
the first/last row plus a few random ellipses / rectangles per image (the geometry rules of
/root/reference/fluid_generator/generate_xiao_fluid_flow.cpp:268-352).  Used by
bench.py, the speed test and the training script when no dataset is given (there is no network)."""
from __future__ import annotations

import numpy as np


def synth_boundary(B, H, W, seed=0):
    """uint8 [B,H,W,1] in {0,1}; image b is drawn from np.random.default_rng(seed + b)."""
    out = np.zeros((B, H, W, 1), dtype=np.uint8)
    sc = min(H / 128.0, 1.0)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        rng = np.random.default_rng(seed + b)
        g = np.zeros((H, W), dtype=bool)
        g[0, :] = True
        g[H - 1, :] = True
        nobj = max(1, int(round(4 * (W / 256.0) * (H / 128.0))))
        for _ in range(nobj):
            kind = rng.integers(0, 2)
            rx = rng.uniform(20, 40) * sc
            ry = rng.uniform(20, 40) * sc
            lo_x, hi_x = 32 * sc + rx, max(W - rx, 32 * sc + rx + 1)
            lo_y, hi_y = 32 * sc, max(H - ry / 2, 32 * sc + 1)
            cx = rng.uniform(lo_x, hi_x)
            cy = rng.uniform(lo_y, hi_y)
            if kind == 0:
                g |= ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0
            else:
                g |= (np.abs(xx - cx) <= rx) & (np.abs(yy - cy) <= ry)
        out[b, :, :, 0] = g
    return out


def synth_sflow(boundary, seed=99):
    """A bounded stand-in target velocity field: 0.1*tanh(N(0,1)), zero inside solids (SURVEY 8d cfg 3)."""
    rng = np.random.default_rng(seed)
    B, H, W, _ = boundary.shape
    t = 0.1 * np.tanh(rng.standard_normal((B, H, W, 2))).astype(np.float32)
    return t * (1.0 - boundary.astype(np.float32))


def load_car_boundaries(npz_path):
    """The 28 bundled car masks (reference cars/*.txt placed and cropped to 128x256), from the
    bit-packed fixture tests/golden/cars_128x256.npz.  Returns uint8 [28,128,256,1]."""
    z = np.load(npz_path)
    shape = tuple(int(s) for s in z["shape"])
    bits = np.unpackbits(z["packed"])[: int(np.prod(shape))]
    return bits.reshape(shape)[..., None].astype(np.uint8)
