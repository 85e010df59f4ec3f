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


def seeded_state_dict(model, seed=0):
    """Deterministic weights for a Resnet18_8s-shaped module, a pure function of
    (parameter name, shape, seed) -- so the golden generator (which builds the REFERENCE
    classes) and the tests (which build ours) get identical tensors without sharing a file.
    Conv weights ~ N(0, sqrt(2/fan_out)); BN gamma ~ U(0.5,1.5), beta ~ N(0,0.1),
    running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5) so folding is exercised."""
    import hashlib

    import torch
    out = {}
    for name, t in model.state_dict().items():
        h = int(hashlib.sha256(f"{seed}:{name}".encode()).hexdigest()[:8], 16)
        g = torch.Generator().manual_seed(h)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros_like(t)
        elif name.endswith("running_var"):
            out[name] = torch.rand(t.shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            out[name] = torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() == 4:
            fan = t.shape[0] * t.shape[2] * t.shape[3]
            out[name] = torch.randn(t.shape, generator=g) * (2.0 / fan) ** 0.5
        elif name.endswith(".weight"):          # BN gamma
            out[name] = torch.rand(t.shape, generator=g) + 0.5
        else:                                    # BN beta / conv bias
            out[name] = torch.randn(t.shape, generator=g) * 0.1
    return out


# ---------------------------------------------------------------- inputs of tests/golden/ref_variants.npz
VARIANT_HWK = (64, 80, 5)


def variant_inputs(seed, n_fg=900, classes=1):
    """mask [1,H,W] int64 with `classes` disc-shaped regions (values 1..classes), vertex
    [1,H,W,K,2] f32: unit vectors towards K planted keypoints, rotated by N(0, 0.05 rad) noise.
    Shared by tests/golden/make_golden_variants.py (which feeds them to the reference's own
    Python functions) and the tests that replay the recorded samples."""
    H, W, K = VARIANT_HWK
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    mask = np.zeros((H, W), np.int64)
    for c in range(classes):
        cx, cy = W * (c + 1) / (classes + 1), H / 2
        d2 = (xx - cx) ** 2 + (yy - cy) ** 2
        order = np.argsort(d2.ravel(), kind="stable")[:n_fg // classes]
        mask.ravel()[order] = c + 1
    kps = np.stack([W / 2 + 30 * np.cos(2 * np.pi * np.arange(K) / K), H / 2 + 20 * np.sin(2 * np.pi * np.arange(K) / K)], 1)
    d = kps[None, None] - np.stack([xx, yy], -1)[:, :, None, :].astype(np.float64)      # [H,W,K,2]
    ang = np.arctan2(d[..., 1], d[..., 0]) + rng.normal(0, 0.05, d.shape[:-1])
    vertex = np.stack([np.cos(ang), np.sin(ang)], -1).astype(np.float32)
    vertex[mask == 0] = 0
    return mask[None], vertex[None], kps
