"""Drop-in for the reference's native extension module ``ransac_voting``
(lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:102-107): same two functions,
same tensor layouts, backed by libpvnet_b200.so.  The reference's own Python wrapper
runs unmodified on top of this module (A/B testing); the fused layer in
``pvnet_b200.ransac_voting_gpu`` does not use it.

    generate_hypothesis(direct[tn,vn,2], coords[tn,2], idxs[hn,vn,2] int32) -> [hn,vn,2]
    voting_for_hypothesis(direct, coords, hypo_pts[hn,vn,2], inliers[hn,vn,tn] uint8, thresh)
    generate_hypothesis_vanishing_point(direct, coords, idxs) -> [hn,vn,3]        (ransac_voting.cpp:61-72)
    voting_for_hypothesis_vanishing_point(direct, coords, hypo_pts[hn,vn,3], inliers, thresh)   (:82-96)
"""
from __future__ import annotations

import ctypes

import torch

from . import _native


def _check(t, name, dtype):
    # CHECK_INPUT of the reference (ransac_voting.cpp:7-9): CUDA + contiguous
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def generate_hypothesis(direct, coords, idxs):
    _check(direct, "direct", torch.float32)
    _check(coords, "coords", torch.float32)
    _check(idxs, "idxs", torch.int32)
    tn, vn, _ = direct.shape
    hn = idxs.shape[0]
    with torch.cuda.device(direct.device):
        hypo = torch.empty([hn, vn, 2], dtype=torch.float32, device=direct.device)
        _native.check(_native.lib().pvnet_generate_hypothesis(
            _p(direct), _p(coords), _p(idxs), _p(hypo), tn, vn, hn,
            ctypes.c_void_p(torch.cuda.current_stream(direct.device).cuda_stream)), "pvnet_generate_hypothesis")
    return hypo


def voting_for_hypothesis(direct, coords, hypo_pts, inliers, inlier_thresh):
    _check(direct, "direct", torch.float32)
    _check(coords, "coords", torch.float32)
    _check(hypo_pts, "hypo_pts", torch.float32)
    _check(inliers, "inliers", torch.uint8)
    tn, vn, _ = direct.shape
    hn = hypo_pts.shape[0]
    with torch.cuda.device(direct.device):
        _native.check(_native.lib().pvnet_voting_for_hypothesis(
            _p(direct), _p(coords), _p(hypo_pts), _p(inliers), tn, vn, hn, float(inlier_thresh),
            ctypes.c_void_p(torch.cuda.current_stream(direct.device).cuda_stream)), "pvnet_voting_for_hypothesis")


def vote_counts(direct, coords, hypo_pts, inlier_thresh):
    """sum_t of the inlier predicate, int32 [hn,vn] (no u8 tensor)."""
    _check(direct, "direct", torch.float32)
    _check(coords, "coords", torch.float32)
    _check(hypo_pts, "hypo_pts", torch.float32)
    tn, vn, _ = direct.shape
    hn = hypo_pts.shape[0]
    with torch.cuda.device(direct.device):
        counts = torch.empty([hn, vn], dtype=torch.int32, device=direct.device)
        _native.check(_native.lib().pvnet_vote_counts(
            _p(direct), _p(coords), _p(hypo_pts), _p(counts), tn, vn, hn, float(inlier_thresh),
            ctypes.c_void_p(torch.cuda.current_stream(direct.device).cuda_stream)), "pvnet_vote_counts")
    return counts


def generate_hypothesis_vanishing_point(direct, coords, idxs):
    _check(direct, "direct", torch.float32)
    _check(coords, "coords", torch.float32)
    _check(idxs, "idxs", torch.int32)
    tn, vn, _ = direct.shape
    hn = idxs.shape[0]
    with torch.cuda.device(direct.device):
        hypo = torch.empty([hn, vn, 3], dtype=torch.float32, device=direct.device)
        _native.check(_native.lib().pvnet_generate_hypothesis_vanishing_point(
            _p(direct), _p(coords), _p(idxs), _p(hypo), tn, vn, hn,
            ctypes.c_void_p(torch.cuda.current_stream(direct.device).cuda_stream)),
            "pvnet_generate_hypothesis_vanishing_point")
    return hypo


def voting_for_hypothesis_vanishing_point(direct, coords, hypo_pts, inliers, inlier_thresh, return_counts=False):
    _check(direct, "direct", torch.float32)
    _check(coords, "coords", torch.float32)
    _check(hypo_pts, "hypo_pts", torch.float32)
    if inliers is not None:
        _check(inliers, "inliers", torch.uint8)
    tn, vn, _ = direct.shape
    hn = hypo_pts.shape[0]
    with torch.cuda.device(direct.device):
        counts = torch.empty([hn, vn], dtype=torch.int32, device=direct.device) if return_counts else None
        _native.check(_native.lib().pvnet_voting_for_hypothesis_vanishing_point(
            _p(direct), _p(coords), _p(hypo_pts), None if inliers is None else _p(inliers),
            None if counts is None else _p(counts), tn, vn, hn, float(inlier_thresh),
            ctypes.c_void_p(torch.cuda.current_stream(direct.device).cuda_stream)),
            "pvnet_voting_for_hypothesis_vanishing_point")
    return counts
