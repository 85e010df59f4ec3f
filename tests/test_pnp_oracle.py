"""CPU: the uncertainty-PnP oracle (oracle/pnp_oracle.py) against the committed fixtures
(tests/golden/pnp_cases.npz, made by tests/golden/make_golden_pnp.py from the reference's data/demo
fixture with OpenCV's P3P / iterative PnP as independent checks) and against first-order optimality."""
import os

import numpy as np
import pytest

from oracle import pnp_oracle as pn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pnp_cases.npz")


@pytest.fixture(scope="module")
def cases():
    return np.load(GOLDEN)


def _names(z):
    return sorted(k[:-5] for k in z.files if k.endswith("_pose"))


def test_known_answer_demo_pose(cases):
    """cat_pose.npy reprojects the object keypoints exactly: any weights must return that pose."""
    for name in ("demo_iso", "demo_aniso"):
        w = pn.covariance_to_weights(cases[name + "_cov"])
        rt = pn.uncertainty_pnp(cases[name + "_kp"], w, cases["points_3d"], cases["K"])
        # keypoints are stored as float32 (1e-5 px rounding): pose recovered to ~1e-6
        assert np.abs(rt - cases[name + "_pose"]).max() < 2e-6, name


def test_minimiser_matches_fixtures_and_is_stationary(cases):
    for name in _names(cases):
        if not name.startswith("noisy"):
            continue
        kp, cov = cases[name + "_kp"], cases[name + "_cov"]
        w = pn.covariance_to_weights(cov)
        rt = pn.uncertainty_pnp(kp, w, cases["points_3d"], cases["K"])
        assert np.abs(rt - cases[name + "_pose"]).max() < 1e-10, name
        x = np.concatenate([pn.rotation_to_rvec(rt[:, :3]), rt[:, 3]])
        args = (kp.astype(np.float64), w, cases["points_3d"], cases["K"])
        g = pn.jacobian(x, *args).T @ pn.residuals(x, *args)
        assert np.abs(g).max() < 1e-8, (name, np.abs(g).max())
        if name + "_cv2iter" in cases.files:          # isotropic noise: OpenCV's own LM reaches the same pose
            assert np.abs(rt - cases[name + "_cv2iter"]).max() < 1e-6, name


def test_weights_closed_form_equals_scipy_sqrtm(cases):
    import scipy.linalg
    cov = cases["noisy_1_cov"].astype(np.float64)
    w = pn.covariance_to_weights(cov)
    for i in range(cov.shape[0]):
        ref = np.linalg.inv(scipy.linalg.sqrtm(cov[i]))          # evaluation_utils.py:176
        assert np.allclose([ref[0, 0], ref[0, 1], ref[1, 1]], w[i], rtol=1e-10, atol=1e-12)
    bad = np.array([[[1e-7, 0], [0, 1.0]], [[np.nan, 0], [0, 1.0]], [[1.0, 2.0], [2.0, 1.0]]])
    assert np.array_equal(pn.covariance_to_weights(bad), np.zeros((3, 3)))


def test_p3p_returns_the_generating_pose():
    rng = np.random.default_rng(3)
    K = np.array([[572.4114, 0., 325.2611], [0., 573.57043, 242.04899], [0., 0., 1.]])
    for _ in range(50):
        P = rng.uniform(-0.1, 0.1, (4, 3))
        R = pn.rodrigues(rng.normal(0, 1, 3))
        t = np.array([rng.uniform(-.2, .2), rng.uniform(-.2, .2), rng.uniform(0.5, 1.5)])
        X = P @ R.T + t
        uv = np.stack([K[0, 0] * X[:, 0] / X[:, 2] + K[0, 2], K[1, 1] * X[:, 1] / X[:, 2] + K[1, 2]], 1)
        got = pn.p3p_init(uv, P, K, np.arange(4))
        assert got is not None and np.abs(got[0] - R).max() < 1e-6 and np.abs(got[1] - t).max() < 1e-6
