"""Device-side uncertainty-driven PnP: the reference's `uncertainty_pnp`
(zju3dv/pvnet lib/utils/extend_utils/extend_utils.py:63-114) and the covariance -> weight step of
`Evaluator.evaluate_uncertainty` (lib/utils/evaluation_utils.py:165-201), over the C ABI
(`pvnet_uncertainty_pnp`, `pvnet_covariance_to_weights` in include/pvnet_b200.h).

    uncertainty_pnp(points_2d [pn,2], weights_2d [pn,3], points_3d [pn,3], camera_matrix [3,3]) -> Rt [3,4]

keeps the reference's signature and return type (numpy float64) for numpy inputs -- the arrays are
moved to the current CUDA device; batched CUDA tensors ([b,pn,2], [b,pn,3]) return a float64 CUDA
tensor [b,3,4] with no host synchronisation, which is what `PoseKeypointPipeline(with_pose=True)` uses so
that poses, not keypoints, are what leaves the GPU.  No CPU path: without the library or a CUDA
device these functions raise.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _camera(camera_matrix):
    k = np.asarray(camera_matrix.detach().cpu() if isinstance(camera_matrix, torch.Tensor) else camera_matrix,
                   dtype=np.float64).reshape(3, 3)
    return (ctypes.c_double * 9)(*k.ravel().tolist())


def covariance_to_weights(cov: torch.Tensor) -> torch.Tensor:
    """cov [...,2,2] CUDA float32 -> weights [...,3] = (wxx, wxy, wyy) of inv(sqrtm(cov))
    (evaluation_utils.py:170-181; zeros where the reference skips the point)."""
    if not cov.is_cuda:
        raise RuntimeError("pvnet_b200: `cov` must be a CUDA tensor (there is no CPU path)")
    c = cov.contiguous().float()
    n = c.numel() // 4
    out = torch.empty(tuple(c.shape[:-2]) + (3,), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        _native.check(_native.lib().pvnet_covariance_to_weights(c.data_ptr(), n, out.data_ptr(), _stream(c.device)),
                      "pvnet_covariance_to_weights")
    return out


def uncertainty_pnp_batched(points_2d, points_3d, camera_matrix, weights_2d=None, cov=None, return_info=False):
    """points_2d [b,pn,2] CUDA; weights_2d [b,pn,3] or cov [b,pn,2,2] (exactly one); points_3d [pn,3];
    camera_matrix 3x3 (host).  -> poses float64 [b,3,4] on the device (and info int32 [b,2])."""
    if not points_2d.is_cuda:
        raise RuntimeError("pvnet_b200: `points_2d` must be a CUDA tensor (there is no CPU path)")
    if (weights_2d is None) == (cov is None):
        raise ValueError("pass exactly one of weights_2d / cov")
    dev = points_2d.device
    p2 = points_2d.contiguous().float()
    b, pn, _ = p2.shape
    p3 = torch.as_tensor(points_3d, dtype=torch.float32, device=dev).contiguous()
    if tuple(p3.shape) != (pn, 3):
        raise ValueError(f"points_3d must be [{pn},3], got {tuple(p3.shape)}")
    w = None if weights_2d is None else weights_2d.to(dev).contiguous().float()
    c = None if cov is None else cov.to(dev).contiguous().float()
    out = torch.empty([b, 3, 4], dtype=torch.float64, device=dev)
    info = torch.empty([b, 2], dtype=torch.int32, device=dev) if return_info else None
    with torch.cuda.device(dev):
        _native.check(_native.lib().pvnet_uncertainty_pnp(
            p2.data_ptr(), None if c is None else c.data_ptr(), None if w is None else w.data_ptr(), p3.data_ptr(),
            _camera(camera_matrix), b, pn, out.data_ptr(), None if info is None else info.data_ptr(), _stream(dev)),
            "pvnet_uncertainty_pnp")
    return (out, info) if return_info else out


def uncertainty_pnp(points_2d, weights_2d, points_3d, camera_matrix):
    """Reference signature (extend_utils.py:63): numpy [pn,2], [pn,3] (wxx,wxy,wyy), [pn,3], [3,3] ->
    Rt numpy float64 [3,4].  Batched CUDA tensors are accepted too (see uncertainty_pnp_batched)."""
    if isinstance(points_2d, torch.Tensor) and points_2d.dim() == 3:
        return uncertainty_pnp_batched(points_2d, points_3d, camera_matrix, weights_2d=weights_2d)
    if not torch.cuda.is_available():
        raise RuntimeError("pvnet_b200: uncertainty_pnp needs a CUDA device (there is no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device())
    p2 = torch.as_tensor(np.asarray(points_2d, np.float32), device=dev)[None]
    w = torch.as_tensor(np.asarray(weights_2d, np.float32), device=dev)[None]
    assert p2.shape[1] == np.asarray(points_3d).shape[0] and p2.shape[1] >= 4          # extend_utils.py:72
    return uncertainty_pnp_batched(p2, np.asarray(points_3d, np.float32), camera_matrix, weights_2d=w)[0].cpu().numpy()
