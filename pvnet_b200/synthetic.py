"""Synthetic inputs for parity tests and bench.py (SURVEY.md §8d / BASELINE.md §5).

NumPy ``default_rng`` only, so the data does not depend on the torch version.
Seeds: ``1000 * config + image_index``.

The shapes mirror what the reference's callers hand to the voting layer
(tools/demo.py:46-55 ``EvalWrapper``): an integer mask [b,h,w] (int64 from
``torch.argmax``) and a vertex field that is a *permuted view* of an NCHW tensor,
i.e. ``ver_pred[b,2K,h,w].permute(0,2,3,1).view(b,h,w,K,2)``.  ``vertex_nchw``
below returns that NCHW array; ``as_reference_view`` does the permute/view.
"""
from __future__ import annotations

import numpy as np

H, W = 480, 640


def disc_mask(n_fg: int, h: int = H, w: int = W, center=(320, 240)) -> np.ndarray:
    """The ``n_fg`` pixels closest to ``center`` (x,y); ties broken by flat index.
    Exact foreground count.  Returns int64 [h,w] in {0,1}."""
    ys, xs = np.mgrid[0:h, 0:w]
    d2 = (xs - center[0]) ** 2 + (ys - center[1]) ** 2
    order = np.lexsort((np.arange(h * w), d2.ravel()))
    m = np.zeros(h * w, np.int64)
    m[order[:n_fg]] = 1
    return m.reshape(h, w)


def random_field(mask: np.ndarray, k: int, seed: int) -> np.ndarray:
    """theta ~ U[0,2pi) per (pixel,keypoint); vertex=(cos,sin) f32, zero outside the
    mask.  Returns NCHW-style [2k,h,w] with channel 2*j = x, 2*j+1 = y of keypoint j."""
    rng = np.random.default_rng(seed)
    h, w = mask.shape
    theta = rng.uniform(0.0, 2.0 * np.pi, size=(k, h, w))
    v = np.empty((2 * k, h, w), np.float32)
    v[0::2] = np.cos(theta)
    v[1::2] = np.sin(theta)
    v *= (mask != 0)[None].astype(np.float32)
    return v


def planted_keypoints(k: int, center=(320.0, 240.0)) -> np.ndarray:
    """kp_j = center + R (cos 2pi j/k, sin 2pi j/k), R=90 (j even) / 260 (j odd)."""
    j = np.arange(k)
    r = np.where(j % 2 == 0, 90.0, 260.0)
    return np.stack([center[0] + r * np.cos(2 * np.pi * j / k),
                     center[1] + r * np.sin(2 * np.pi * j / k)], 1)


def planted_field(mask: np.ndarray, k: int, seed: int, sigma: float = 0.03):
    """Unit vectors from each pixel towards kp_j, rotated by eps ~ N(0, sigma rad).
    Same normalisation rule as the reference's demo recipe (tools/demo.py:64-67:
    norms below 1e-3 get +1e-3).  Returns ([2k,h,w] f32, keypoints [k,2] f64)."""
    rng = np.random.default_rng(seed)
    h, w = mask.shape
    kps = planted_keypoints(k)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    v = np.zeros((2 * k, h, w), np.float32)
    for j in range(k):
        dx, dy = kps[j, 0] - xs, kps[j, 1] - ys
        n = np.sqrt(dx * dx + dy * dy)
        n[n < 1e-3] += 1e-3
        dx, dy = dx / n, dy / n
        eps = rng.normal(0.0, sigma, size=(h, w)) if sigma > 0 else np.zeros((h, w))
        c, s = np.cos(eps), np.sin(eps)
        v[2 * j] = (c * dx - s * dy) * (mask != 0)
        v[2 * j + 1] = (s * dx + c * dy) * (mask != 0)
    return v, kps


def as_reference_view(vertex_nchw: np.ndarray) -> np.ndarray:
    """[b,2k,h,w] -> the [b,h,w,k,2] strided view EvalWrapper builds (demo.py:48-50)."""
    b, c2, h, w = vertex_nchw.shape
    s = vertex_nchw.strides
    return np.lib.stride_tricks.as_strided(vertex_nchw, shape=(b, h, w, c2 // 2, 2),
                                           strides=(s[0], s[2], s[3], 2 * s[1], s[1]), writeable=False)


def draw_idxs(tn: int, hn: int, k: int, seed: int, rounds: int | None = None) -> np.ndarray:
    """Injected pixel-pair indices: int32 [hn,k,2] (or [rounds,hn,k,2]) in [0,tn)."""
    rng = np.random.default_rng(seed)
    shape = (hn, k, 2) if rounds is None else (rounds, hn, k, 2)
    return rng.integers(0, max(tn, 1), size=shape, dtype=np.int32)


def selection_field(seed: int, h: int = H, w: int = W) -> np.ndarray:
    """Stand-in for the reference's `uniform_(0,1)` subsample field (float32 [h,w])."""
    return np.random.default_rng(seed).random((h, w), dtype=np.float32)


def backbone_input(b: int, seed: int, h: int = H, w: int = W) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((b, 3, h, w), dtype=np.float32)
