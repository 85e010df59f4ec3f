"""CPU: the oracle's restatements of the module's other voting variants against
tests/golden/ref_variants.npz -- outputs of the REFERENCE'S OWN Python functions
(ransac_voting_gpu.py:10, :218, :263, :960) run on the CPU with the compiled extension stubbed
by the oracle kernels (tests/golden/make_golden_variants.py), replaying the samples the
reference drew."""
import os

import numpy as np
import pytest

from oracle import pvnet_oracle as po
from tests.helpers import variant_inputs


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_variants.npz"))


def _no_tie_at_cut(ratio, topk):
    """Inlier ratios are count/tn, so equal values are common; when the k-th and (k+1)-th largest
    are equal, which hypothesis torch.topk keeps is unspecified -> compare only the other keypoints."""
    srt = -np.sort(-ratio, axis=1)
    return srt[:, topk - 1] != srt[:, topk]


def test_v1_winner_hypothesis_is_bit_exact(gold):
    seed, n, classes = gold["v1_seed"]
    mask, vertex, _ = variant_inputs(int(seed), int(n), int(classes))
    idxs = [[gold["v1_idxs"][k] for k in range(classes)]]
    out = po.ransac_voting_layer(mask, vertex, classes + 1, 32, inlier_thresh=0.99, idxs=idxs)
    assert out.shape == gold["v1_out"].shape
    assert np.array_equal(out, gold["v1_out"])          # a hypothesis is pure kernel output: no tolerance


def test_ransac_voting_hypothesis_bit_exact(gold):
    seed, n, _ = gold["hyp_seed"]
    mask, vertex, _ = variant_inputs(int(seed), int(n))
    hyp, cnt = po.ransac_voting_hypothesis(mask, vertex, 48, inlier_thresh=0.99, idxs=[gold["hyp_idxs"]])
    assert np.array_equal(hyp, gold["hyp_out"])
    assert cnt.dtype == np.int64 and np.array_equal(cnt, gold["hyp_counts"])


def test_ransac_voting_hypothesis_skip_branch():
    mask, vertex, _ = variant_inputs(5, 3)               # 3 foreground pixels < min_num
    hyp, cnt = po.ransac_voting_hypothesis(mask, vertex, 8, idxs=[None])
    assert not hyp.any() and (cnt == 1).all()            # ransac_voting_gpu.py:228-233


def test_estimate_voting_distribution(gold):
    seed, n, _ = gold["dist_seed"]
    mask, vertex, _ = variant_inputs(int(seed), int(n))
    mean, cov, ratio = po.estimate_voting_distribution(mask, vertex, 32, 96, 24, inlier_thresh=0.99,
                                                       idxs=[gold["dist_idxs"]], return_ratio=True)
    ok = _no_tie_at_cut(ratio[0], 24)
    assert ok.sum() >= 2
    # the reference sums 24 weighted points in fp32; the restatement in fp64
    assert np.abs(mean - gold["dist_mean"])[0, ok].max() <= 1e-4 * 80
    assert np.abs(cov - gold["dist_cov"])[0, ok].max() <= 1e-4 * max(1.0, np.abs(gold["dist_cov"]).max())


def test_ransac_motion_voting(gold):
    if "motion_out" not in gold.files:
        pytest.skip("the reference's ransac_motion_voting did not run under the torch that made the fixture")
    seed, n, _ = gold["motion_seed"]
    mask, vertex, _ = variant_inputs(int(seed), int(n))
    out = po.ransac_motion_voting(mask, vertex)
    assert np.abs(out - gold["motion_out"]).max() <= 1e-4      # fp32 mean of 500 values near 40
    empty = po.ransac_motion_voting(np.zeros_like(mask), vertex)
    assert not empty.any()


def test_v4_variance_matches_direct_residuals():
    """v4's var (:750-752) against an independent evaluation from the debug output of v3."""
    mask, vertex, _ = variant_inputs(21, 600)
    rng = np.random.default_rng(0)
    idxs = [rng.integers(0, 600, (40, 5, 2), dtype=np.int32)]
    kp, var = po.ransac_voting_layer_v4(mask, vertex, 40, idxs=idxs)
    kp3, dbg = po.ransac_voting_layer_v3(mask, vertex, 40, inlier_thresh=0.99, idxs=idxs, return_debug=True)
    assert np.array_equal(kp, kp3)
    d = dbg[0]
    for k in range(5):
        sel = d["refit_inliers"][k].astype(bool)
        n = np.stack([d["direct"][sel, k, 1], -d["direct"][sel, k, 0]], 1).astype(np.float64)
        r = n @ kp[0, k].astype(np.float64) - (n * d["coords"][sel]).sum(1)
        assert abs(var[0, k] - (r * r).mean()) <= 1e-5 * max(1.0, (r * r).mean())
    few, var_few = po.ransac_voting_layer_v4(*variant_inputs(5, 3)[:2], 8, idxs=[None])
    assert not few.any() and (var_few == 1).all()        # :685-689
