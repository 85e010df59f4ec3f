"""CPU tests: the oracle against the golden fixtures and its own invariants."""
import os

import numpy as np
import pytest

from oracle import pvnet_oracle as po
from pvnet_b200 import synthetic as syn
from tests.helpers import GOLDEN, cfg1_inputs, demo_fixture


def test_demo_known_answer():
    """voting on compute_vertex(mask, points_2d) returns points_2d."""
    mask, field, pts = demo_fixture()
    assert int(mask.sum()) == 2289
    vertex = syn.as_reference_view(field[None])
    idxs = syn.draw_idxs(2289, 512, 9, seed=7)
    kp = po.ransac_voting_layer_v3(mask[None], vertex, 512, inlier_thresh=0.99, idxs=[idxs])
    assert kp.shape == (1, 9, 2)
    assert np.abs(kp[0] - pts).max() < 1e-2


@pytest.mark.parametrize("kind", ["random", "planted"])
def test_frozen_cfg1(kind):
    z = np.load(os.path.join(GOLDEN, "oracle_cfg1.npz"))
    mask, field, idxs = cfg1_inputs(kind)
    kp, dbg = po.ransac_voting_layer_v3(mask[None], syn.as_reference_view(field[None]), 128,
                                        inlier_thresh=0.99, idxs=[idxs], return_debug=True)
    assert np.array_equal(dbg[0]["counts"], z[kind + "_counts"])
    assert np.array_equal(dbg[0]["win_idx"], z[kind + "_win_idx"])
    assert np.float64(dbg[0]["hyp"].astype(np.float64).sum()) == z[kind + "_hyp_sum"]
    np.testing.assert_array_equal(kp[0], z[kind + "_kp"])


def test_counts_equal_sum_of_inlier_tensor():
    mask, field, idxs = cfg1_inputs("planted")
    coords, direct = po.compact(mask.astype(np.uint8), syn.as_reference_view(field[None])[0])
    hyp = po.generate_hypothesis_kernel(direct, coords, idxs[:16])
    inl = po.voting_for_hypothesis_kernel(direct, coords, hyp, 0.99)
    assert inl.shape == (16, 9, 10000) and inl.max() == 1
    assert np.array_equal(inl.sum(2).astype(np.int32), po.vote_counts(direct, coords, hyp, 0.99))


def test_degenerate_pair_stays_zero_and_is_scored():
    """ransac_voting_kernel.cu:42-43,75: parallel rays leave hypothesis (0,0); it is
    still voted on."""
    direct = np.zeros((4, 1, 2), np.float32)
    direct[:, 0] = [1.0, 0.0]
    coords = np.array([[5, 5], [9, 5], [5, 9], [1, 1]], np.float32)
    hyp = po.generate_hypothesis_kernel(direct, coords, np.array([[[0, 1]], [[2, 2]]], np.int32))
    assert np.array_equal(hyp, np.zeros((2, 1, 2), np.float32))
    # pixel (1,1) with direction (1,0) does not point at (0,0); craft one that does
    direct[3, 0] = [-np.sqrt(0.5), -np.sqrt(0.5)]
    assert po.vote_counts(direct, coords, hyp, 0.99)[0, 0] == 1


def test_hypothesis_on_pixel_is_not_inlier():
    """norm2 < 1e-6 -> skip (ransac_voting_kernel.cu:121)."""
    direct = np.array([[[1.0, 0.0]]], np.float32)
    coords = np.array([[3.0, 4.0]], np.float32)
    assert po.vote_counts(direct, coords, np.array([[[3.0, 4.0]]], np.float32), 0.5)[0, 0] == 0
    assert po.vote_counts(direct, coords, np.array([[[4.0, 4.0]]], np.float32), 0.5)[0, 0] == 1
    zero_dir = np.zeros((1, 1, 2), np.float32)
    assert po.vote_counts(zero_dir, coords, np.array([[[4.0, 4.0]]], np.float32), -1.0)[0, 0] == 0


def test_skip_and_first_max():
    mask = np.zeros((1, 16, 16), np.int64)
    mask[0, 0, :4] = 1      # 4 < min_num=5
    vertex = np.zeros((1, 16, 16, 2, 2), np.float32)
    assert np.array_equal(po.ransac_voting_layer_v3(mask, vertex, 8, idxs=[None]), np.zeros((1, 2, 2), np.float32))
    # two identical hypotheses: winner is the lower index
    counts = np.array([[3, 1], [3, 5], [2, 5]])
    assert counts.argmax(0).tolist() == [0, 1]


def test_mask_semantics_v3_nonzero_vs_cov_equals_one():
    """v3: nonzero after .byte() (:527); with_mean: == 1 (:339)."""
    assert po._byte_mask(np.array([0, 1, 2, 256, 257], np.int64)).tolist() == [0, 1, 2, 0, 1]


def test_subsample_matches_reference_formula():
    mask = syn.disc_mask(40000)
    field = syn.planted_field(mask, 3, 5, sigma=0.0)[0]
    sel = syn.selection_field(11)
    p = po.subsample_probability(30000, 40000)
    kept = int(((sel < p) & (mask != 0)).sum())
    idxs = syn.draw_idxs(kept, 64, 3, seed=3)
    kp, dbg = po.ransac_voting_layer_v3(mask[None], syn.as_reference_view(field[None]), 64, inlier_thresh=0.99,
                                        max_num=30000, idxs=[idxs], selection=[sel], return_debug=True)
    assert dbg[0]["tn"] == kept and 29000 < kept < 31000
    assert np.abs(kp[0] - syn.planted_keypoints(3)).max() < 1e-2


def test_covariance_with_mean_small():
    mask = syn.disc_mask(3000)
    field, kps = syn.planted_field(mask, 4, 21, sigma=0.05)
    vertex = syn.as_reference_view(field[None])
    idxs_v3 = syn.draw_idxs(3000, 64, 4, seed=1)
    mean = po.ransac_voting_layer_v3(mask[None], vertex, 64, inlier_thresh=0.99, idxs=[idxs_v3])
    idxs = syn.draw_idxs(3000, 32, 4, seed=2, rounds=4)
    m2, cov = po.estimate_voting_distribution_with_mean(mask[None], vertex, mean, round_hyp_num=32, min_hyp_num=128,
                                                        inlier_thresh=0.99, idxs=[idxs])
    assert m2 is not None and cov.shape == (1, 4, 2, 2)
    assert np.allclose(cov[0, :, 0, 1], cov[0, :, 1, 0])
    ev = np.linalg.eigvalsh(cov[0].astype(np.float64))
    assert (ev > -1e-6).all()
    # far keypoints (odd index, R=260) are less certain than near ones (R=90)
    assert np.trace(cov[0, 1]) > np.trace(cov[0, 0])


def test_covariance_skip_branch():
    mask = np.zeros((1, 8, 8), np.int64)
    vertex = np.zeros((1, 8, 8, 2, 2), np.float32)
    mean = np.array([[[1.0, 2.0], [0.0, 0.0]]], np.float32)
    _, cov = po.estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=8, min_hyp_num=16,
                                                       idxs=[None])
    expect = np.outer(mean[0, 0], mean[0, 0]) * 16 / (16 + 1e-3)
    assert np.allclose(cov[0, 0], expect, rtol=1e-6)
    assert np.array_equal(cov[0, 1], np.zeros((2, 2), np.float32))


def test_v5_confidence_oracle():
    mask = syn.disc_mask(2000)
    field = syn.planted_field(mask, 3, 8, sigma=0.0)[0]
    idxs = syn.draw_idxs(2000, 64, 3, seed=2)
    kp, conf = po.ransac_voting_layer_v5(mask[None], syn.as_reference_view(field[None]), 64, inlier_thresh=0.99,
                                         max_num=30000, idxs=[idxs])
    assert kp.shape == (1, 3, 2) and conf.shape == (1, 3)
    assert (conf > 0.95).all()       # exact field: nearly every pixel agrees with the refitted point at 0.999
