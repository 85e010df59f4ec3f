"""GPU: pvnet_uncertainty_pnp / pvnet_covariance_to_weights against the oracle and the fixtures.
Bar: the device pose equals the minimiser of the reference's cost (uncertainty_pnp.cpp:20-37) reached
from the reference's start -- 1e-8 on R and t against the fp64 oracle (inputs are the same float32
keypoints / covariances on both sides)."""
import os

import numpy as np
import pytest
import torch

from oracle import pnp_oracle as pn
from pvnet_b200 import extend_utils as eu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pnp_cases.npz")


def test_batched_device_pnp_matches_fixtures_and_oracle():
    z = np.load(GOLDEN)
    names = sorted(k[:-5] for k in z.files if k.endswith("_pose"))
    kp = torch.from_numpy(np.stack([z[n + "_kp"] for n in names])).to(DEV)
    cov = torch.from_numpy(np.stack([z[n + "_cov"] for n in names])).to(DEV)
    poses, info = eu.uncertainty_pnp_batched(kp, z["points_3d"], z["K"], cov=cov, return_info=True)
    poses, info = poses.cpu().numpy(), info.cpu().numpy()
    assert (info[:, 0] == 0).all(), info
    pts32 = z["points_3d"].astype(np.float32)
    for i, n in enumerate(names):
        w = pn.covariance_to_weights(z[n + "_cov"])
        ref = pn.uncertainty_pnp(z[n + "_kp"], w, pts32, z["K"])        # the device reads float32 object points
        assert np.abs(poses[i] - ref).max() < 1e-8, (n, np.abs(poses[i] - ref).max(), info[i])
        assert np.abs(poses[i] - z[n + "_pose"]).max() < 2e-6, n           # fixture (float64 object points)
        R = poses[i][:, :3]
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(R) - 1) < 1e-12


def test_weights_and_reference_signature():
    z = np.load(GOLDEN)
    cov = z["noisy_4_cov"]
    w_dev = eu.covariance_to_weights(torch.from_numpy(cov).to(DEV)).cpu().numpy()
    assert np.allclose(w_dev, pn.covariance_to_weights(cov), rtol=2e-6, atol=1e-7)
    bad = torch.tensor([[[1e-7, 0], [0, 1.0]], [[float("nan"), 0], [0, 1.0]], [[1.0, 2.0], [2.0, 1.0]]], device=DEV)
    assert torch.equal(eu.covariance_to_weights(bad), torch.zeros(3, 3, device=DEV))
    # reference call shape: numpy in, numpy [3,4] float64 out (extend_utils.py:63-114)
    w = pn.covariance_to_weights(cov)
    rt = eu.uncertainty_pnp(z["noisy_4_kp"], w, z["points_3d"], z["K"])
    assert isinstance(rt, np.ndarray) and rt.shape == (3, 4) and rt.dtype == np.float64
    ref = pn.uncertainty_pnp(z["noisy_4_kp"], w.astype(np.float32), z["points_3d"].astype(np.float32), z["K"])
    assert np.abs(rt - ref).max() < 1e-8
    # pn == 4: the P3P pose itself (extend_utils.py:90-94)
    idx = np.argsort(w[:, 0] + w[:, 1], kind="stable")[-4:]
    rt4 = eu.uncertainty_pnp(z["demo_iso_kp"][idx], w[idx], z["points_3d"][idx], z["K"])
    assert np.abs(rt4 - z["demo_iso_pose"]).max() < 1e-4


def test_random_poses_many_images():
    rng = np.random.default_rng(9)
    z = np.load(GOLDEN)
    pts, K = z["points_3d"].astype(np.float32), z["K"]
    b = 64
    kps, covs, truth = [], [], []
    for _ in range(b):
        R = pn.rodrigues(rng.normal(0, 1.0, 3))
        t = np.array([rng.uniform(-.15, .15), rng.uniform(-.15, .15), rng.uniform(0.5, 1.5)])
        X = pts @ R.T + t
        uv = np.stack([K[0, 0] * X[:, 0] / X[:, 2] + K[0, 2], K[1, 1] * X[:, 1] / X[:, 2] + K[1, 2]], 1)
        A = rng.normal(0, 1, (9, 2, 2))
        cov = A @ A.transpose(0, 2, 1) + 0.2 * np.eye(2)
        kps.append(uv + np.stack([rng.multivariate_normal(np.zeros(2), c) for c in cov]))
        covs.append(cov)
        truth.append(np.concatenate([R, t[:, None]], 1))
    kp = torch.from_numpy(np.stack(kps).astype(np.float32)).to(DEV)
    cov = torch.from_numpy(np.stack(covs).astype(np.float32)).to(DEV)
    poses, info = eu.uncertainty_pnp_batched(kp, pts, K, cov=cov, return_info=True)
    poses, info = poses.cpu().numpy(), info.cpu().numpy()
    assert (info[:, 0] & 2 == 0).all()
    worst = 0.0
    for i in range(b):
        ref = pn.uncertainty_pnp(kp[i].cpu().numpy(), pn.covariance_to_weights(cov[i].cpu().numpy()), pts, K)
        worst = max(worst, np.abs(poses[i] - ref).max())
        assert np.abs(poses[i][:, 3] - truth[i][:, 3]).max() < 0.3          # sane: near the generating pose (depth is the weak axis)
    assert worst < 1e-8, worst
