"""Shared test helpers (CPU side)."""
import os

import numpy as np

from pvnet_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def demo_fixture():
    """mask int64 [480,640], exact vector field NCHW [18,480,640] f32, points_2d [9,2].
    The field is rebuilt from the fixture with the recipe of the reference's
    tools/demo.py:58-71 (`compute_vertex`): (kp - xy)/norm, norm<1e-3 -> +1e-3."""
    z = np.load(os.path.join(GOLDEN, "demo_cat.npz"))
    h, w = (int(v) for v in z["shape"])
    fg = z["fg_yx"].astype(np.int64)
    pts = z["points_2d"]
    mask = np.zeros((h, w), np.int64)
    mask[fg[:, 0], fg[:, 1]] = 1
    xy = fg[:, [1, 0]].astype(np.float64)
    v = pts[None, :, :2] - xy[:, None, :]
    n = np.linalg.norm(v, axis=2, keepdims=True)
    n[n < 1e-3] += 1e-3
    v = v / n
    field = np.zeros((h, w, pts.shape[0], 2), np.float32)
    field[fg[:, 0], fg[:, 1]] = v
    nchw = np.ascontiguousarray(field.reshape(h, w, -1).transpose(2, 0, 1))
    return mask, nchw, pts


def cfg1_inputs(kind):
    mask = syn.disc_mask(10000)
    field = syn.random_field(mask, 9, 1000) if kind == "random" else syn.planted_field(mask, 9, 1000)[0]
    idxs = syn.draw_idxs(10000, 128, 9, seed=1000)
    return mask, field, idxs
