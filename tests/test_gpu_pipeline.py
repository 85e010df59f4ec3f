"""GPU parity tests of round 2's voting path: the fused v3 + covariance pipeline
(`pvnet_ransac_voting_pipeline`) at BASELINE configs 4 and 5's exact shapes, config 3's 12-point
sweep, the device-side sampler, and an adversarial test of k_vote3's guard band.

Bar (same as tests/test_gpu_vote.py): hypotheses and inlier counts bit-exact against the oracle /
the exact-sequence kernel; keypoints within 1e-4; covariances atol 1e-4 + rtol 1e-4.
"""
import numpy as np
import pytest
import torch

from oracle import pvnet_oracle as po
from pvnet_b200 import ransac_voting as ext
from pvnet_b200 import ransac_voting_gpu as rv
from pvnet_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dev(mask_np, field_np, mask_dtype=torch.int64):
    mask = torch.from_numpy(np.ascontiguousarray(mask_np)).to(DEV).to(mask_dtype)
    ver = torch.from_numpy(np.ascontiguousarray(field_np)).to(DEV)
    b, c2, h, w = ver.shape
    return mask, ver.permute(0, 2, 3, 1).view(b, h, w, c2 // 2, 2)


def _check_pipeline(masks, fields, hn, cov_hn, cov_min, thresh, cov_thresh, max_num, sels=None, seed=0,
                    mask_dtype=torch.int64, pixel_major=False):
    """Fused call with injected samples vs oracle v3 -> oracle with_mean(mean = oracle v3)."""
    b, vn = masks.shape[0], fields.shape[1] // 2
    rounds = -(-cov_min // cov_hn)
    hnt = rounds * cov_hn
    rng = np.random.default_rng(seed)
    idxs = rng.integers(0, 2 ** 31 - 1, (b, hn, vn, 2), dtype=np.int32)
    cov_idxs = rng.integers(0, 2 ** 31 - 1, (b, hnt, vn, 2), dtype=np.int32)
    mask, vertex = _dev(masks, fields, mask_dtype)
    if pixel_major:
        vertex = vertex.contiguous()
    sel = None if sels is None else torch.from_numpy(np.stack(sels)).to(DEV)
    kp, cov, dbg = rv.ransac_voting_pipeline(mask, vertex, hn, thresh, True, cov_hn, cov_min, cov_thresh,
                                             max_num=max_num, idxs=torch.from_numpy(idxs),
                                             cov_idxs=torch.from_numpy(cov_idxs), selection=sel, rng="none",
                                             return_debug=True)
    torch.cuda.synchronize()
    tn = dbg["tn"].cpu().numpy()
    view = syn.as_reference_view(fields)
    # the oracle takes samples already reduced modulo tn (the product reduces them itself)
    o_idxs = [idxs[i] % max(int(tn[i]), 1) for i in range(b)]
    o_cidx = [(cov_idxs[i] % max(int(tn[i]), 1)).reshape(rounds, cov_hn, vn, 2) for i in range(b)]
    okp, odbg = po.ransac_voting_layer_v3(masks, view, hn, inlier_thresh=thresh, max_num=max_num, idxs=o_idxs,
                                          selection=sels, return_debug=True)
    _, ocov, ocdbg = po.estimate_voting_distribution_with_mean(masks, view, okp, cov_hn, cov_min, inlier_thresh=cov_thresh,
                                                               max_num=max_num, idxs=o_cidx, selection=sels,
                                                               return_debug=True)
    for bi in range(b):
        assert tn[bi] == odbg[bi]["tn"] == ocdbg[bi]["tn"]
        assert np.array_equal(dbg["hyp"][bi].cpu().numpy().view(np.uint32), odbg[bi]["hyp"].view(np.uint32))
        assert np.array_equal(dbg["counts"][bi].cpu().numpy(), odbg[bi]["counts"]), "v3 counts not bit-exact"
        assert np.array_equal(dbg["cov_hyp"][bi].cpu().numpy().view(np.uint32), ocdbg[bi]["hyp"].view(np.uint32))
        assert np.array_equal(dbg["cov_counts"][bi].cpu().numpy(), ocdbg[bi]["counts"]), "cov counts not bit-exact"
    assert np.abs(kp.cpu().numpy() - okp).max() <= 1e-4
    # the product's covariance uses ITS mean (within 1e-4 of the oracle's): compare at the same tolerance as elsewhere
    assert np.allclose(cov.cpu().numpy(), ocov, atol=1e-4 + 2e-4 * np.abs(ocov).max() ** 0.5, rtol=1e-4), \
        np.abs(cov.cpu().numpy() - ocov).max()
    return kp, cov, dbg


def test_config4_shape_fused_vs_oracle():
    """BASELINE config 4 per image: K=9, 20000 px, v3(256) + with_mean(256, 4096), thresh 0.99."""
    masks = np.stack([syn.disc_mask(20000), syn.disc_mask(20000, center=(300, 200))])
    fields = np.stack([syn.planted_field(masks[i], 9, 4000 + i, sigma=0.05)[0] for i in range(2)])
    _check_pipeline(masks, fields, 256, 256, 4096, 0.99, 0.99, 30000, seed=4)


def test_config4_shape_random_field_u8_mask_pixel_major():
    """Same sizes on the random field (every hypothesis scores ~4.5 % inliers: no pruning possible), a
    uint8 mask and a pixel-major (contiguous [b,h,w,K,2]) field: the layouts the fused backbone emits."""
    masks = np.stack([syn.disc_mask(20000)])
    fields = np.stack([syn.random_field(masks[0], 9, 4100)])
    _check_pipeline(masks, fields, 256, 256, 4096, 0.99, 0.99, 30000, seed=5, mask_dtype=torch.uint8,
                    pixel_major=True)


def test_config5_shape_fused_vs_oracle():
    """BASELINE config 5 per image: K=17, 92160 px (30 %), v3(1024, max_num=30000) with an injected
    selection field (subsampling active) + with_mean(1024, 1024)."""
    masks = np.stack([syn.disc_mask(92160)])
    fields = np.stack([syn.planted_field(masks[0], 17, 5000, sigma=0.05)[0]])
    _check_pipeline(masks, fields, 1024, 1024, 1024, 0.99, 0.99, 30000, sels=[syn.selection_field(51)], seed=6)


def test_pipeline_different_thresholds_two_launches():
    masks = np.stack([syn.disc_mask(6000)])
    fields = np.stack([syn.planted_field(masks[0], 5, 77, sigma=0.1)[0]])
    _check_pipeline(masks, fields, 128, 64, 200, 0.999, 0.99, 30000, seed=8)


def _compat_lists(mask_np, field_np):
    coords, direct = po.compact(mask_np.astype(np.uint8), syn.as_reference_view(field_np[None])[0])
    return torch.from_numpy(direct).to(DEV), torch.from_numpy(coords).to(DEV)


@pytest.mark.parametrize("n_fg", [1000, 10000, 50000, 150000])
@pytest.mark.parametrize("hn", [128, 512, 2048])
def test_config3_sweep_counts_vs_exact_kernel(n_fg, hn):
    """BASELINE config 3 (max_num = 10**9 so nothing is subsampled): the fast kernel's counts equal the
    exact-sequence kernel's (`pvnet_vote_counts`, the reference's instruction sequence per test)."""
    mask_np = syn.disc_mask(n_fg)
    field = syn.random_field(mask_np, 9, 3000 + n_fg % 97)
    mask, vertex = _dev(mask_np[None], field[None])
    idxs = torch.from_numpy(syn.draw_idxs(n_fg, hn, 9, seed=hn)[None])
    kp, dbg = rv.ransac_voting_layer_v3(mask, vertex, hn, inlier_thresh=0.99, max_num=10 ** 9, idxs=idxs,
                                        return_debug=True)
    d, c = _compat_lists(mask_np, field)
    exact = ext.vote_counts(d, c, dbg["hyp"][0].contiguous(), 0.99)
    assert int(dbg["tn"][0]) == n_fg
    assert torch.equal(dbg["counts"][0], exact)
    assert torch.isfinite(kp).all()


@pytest.mark.parametrize("T", [0.99, 0.999, 0.9, 0.5])
def test_guard_band_adversarial(T):
    """Every pixel's direction sits within +-4e-6 (relative, in angle) of the cone edge of a hypothesis
    that the sampled pairs reproduce: nearly every test is inside or next to k_vote3's guard band."""
    rng = np.random.default_rng(11)
    th = np.arccos(float(np.float32(T)))
    mask_np = syn.disc_mask(6000)
    H = np.array([[401.37, 163.91], [95.03, 402.2], [330.11, 250.77]])          # one target per keypoint
    ys, xs = np.mgrid[0:480, 0:640].astype(np.float64)
    field = np.zeros((6, 480, 640), np.float32)
    for k in range(3):
        ang = np.arctan2(H[k, 1] - ys, H[k, 0] - xs)
        delta = th * (1.0 + rng.uniform(-4e-6, 4e-6, ang.shape)) * rng.choice([-1.0, 1.0], ang.shape)
        anchor = rng.random(ang.shape) < 0.02             # 2 % of the pixels point exactly at H: the samples
        a = np.where(anchor, ang, ang + delta)
        field[2 * k] = np.cos(a) * mask_np
        field[2 * k + 1] = np.sin(a) * mask_np
    coords, direct = po.compact(mask_np.astype(np.uint8), syn.as_reference_view(field[None])[0])
    hn = 64
    idxs = np.zeros((hn, 3, 2), np.int32)                 # pairs of anchor pixels per keypoint
    for k in range(3):
        ang = np.arctan2(H[k, 1] - coords[:, 1], H[k, 0] - coords[:, 0])
        is_anchor = np.abs(np.arctan2(direct[:, k, 1], direct[:, k, 0]) - ang) < 1e-4
        idxs[:, k, :] = rng.choice(np.nonzero(is_anchor)[0], (hn, 2))
    mask, vertex = _dev(mask_np[None], field[None])
    kp, dbg = rv.ransac_voting_layer_v3(mask, vertex, hn, inlier_thresh=T, idxs=torch.from_numpy(idxs[None]),
                                        return_debug=True)
    ohyp = po.generate_hypothesis_kernel(direct, coords, idxs)
    assert np.array_equal(dbg["hyp"][0].cpu().numpy().view(np.uint32), ohyp.view(np.uint32))
    ocnt = po.vote_counts(direct, coords, ohyp, T)
    cnt = dbg["counts"][0].cpu().numpy()
    assert np.array_equal(cnt, ocnt), np.abs(cnt - ocnt).max()
    # the test is only adversarial if the counts are neither ~0 nor ~tn
    assert 0.1 * 6000 < np.median(ocnt) < 0.9 * 6000


def test_huge_and_degenerate_hypotheses_far_tiles():
    """Hypotheses at ~1e5..1e7 px (near-parallel pairs) and a mask split in two far-apart blobs (tile
    bounding boxes as wide as the image)."""
    m = syn.disc_mask(1500, center=(40, 40)) | syn.disc_mask(1500, center=(600, 440))
    rng = np.random.default_rng(2)
    base = rng.uniform(0, 2 * np.pi, 4)
    ys, xs = np.mgrid[0:480, 0:640]
    field = np.zeros((8, 480, 640), np.float32)
    for k in range(4):
        a = base[k] + rng.normal(0, 1e-4, (480, 640))          # nearly parallel rays: far intersections
        field[2 * k], field[2 * k + 1] = np.cos(a) * m, np.sin(a) * m
    mask, vertex = _dev(m[None], field[None])
    idxs = torch.from_numpy(syn.draw_idxs(3000, 200, 4, seed=3)[None])
    kp, dbg = rv.ransac_voting_layer_v3(mask, vertex, 200, inlier_thresh=0.99, idxs=idxs, return_debug=True)
    d, c = _compat_lists(m, field)
    assert torch.equal(dbg["counts"][0], ext.vote_counts(d, c, dbg["hyp"][0].contiguous(), 0.99))
    assert float(dbg["hyp"].abs().max()) > 1e4


def test_device_rng_reproducible_fresh_and_uniform():
    masks = np.stack([syn.disc_mask(9000), syn.disc_mask(50000)])
    fields = np.stack([syn.planted_field(masks[i], 9, 900 + i)[0] for i in range(2)])
    mask, vertex = _dev(masks, fields)
    torch.manual_seed(123)
    rv.reset_device_rng(DEV)
    kp1, cov1, d1 = rv.ransac_voting_pipeline(mask, vertex, 256, 0.99, True, 128, 512, return_debug=True)
    kp2, cov2, d2 = rv.ransac_voting_pipeline(mask, vertex, 256, 0.99, True, 128, 512, return_debug=True)
    rv.reset_device_rng(DEV)
    kp3, cov3, d3 = rv.ransac_voting_pipeline(mask, vertex, 256, 0.99, True, 128, 512, return_debug=True)
    assert torch.equal(d1["hyp"], d3["hyp"]) and torch.equal(kp1, kp3) and torch.equal(cov1, cov3)   # same seed
    assert not torch.equal(d1["hyp"], d2["hyp"])                                                      # offset advanced
    # image 1 has 50000 > 30000 foreground pixels: the device-side Bernoulli keeps about max_num of them
    tn = d1["tn"].cpu().numpy()
    assert tn[0] == 9000 and abs(int(tn[1]) - 30000) < 6 * (30000 * 0.4) ** 0.5
    assert int(d2["tn"][1]) != int(tn[1])
    # results agree with the planted keypoints (the sampler feeds a working RANSAC)
    kps = syn.planted_keypoints(9)
    assert np.abs(kp1.cpu().numpy() - kps[None])[:, 0::2].max() < 3.0      # the keypoints near the object (R = 90 px)
    assert torch.isfinite(cov1).all()


def test_device_rng_matches_injected_replay():
    """The hypotheses the device sampler produced are ray intersections of pixel pairs of the list:
    counts recomputed by the exact kernel on those hypotheses agree."""
    mask_np = syn.disc_mask(7000)
    field = syn.planted_field(mask_np, 9, 31)[0]
    mask, vertex = _dev(mask_np[None], field[None])
    kp, dbg = rv.ransac_voting_layer_v3(mask, vertex, 300, inlier_thresh=0.99, rng="device", return_debug=True)
    d, c = _compat_lists(mask_np, field)
    assert torch.equal(dbg["counts"][0], ext.vote_counts(d, c, dbg["hyp"][0].contiguous(), 0.99))


def test_workspace_is_reused_between_calls():
    mask_np = syn.disc_mask(3000)
    field = syn.planted_field(mask_np, 9, 1)[0]
    mask, vertex = _dev(mask_np[None], field[None])
    rv.ransac_voting_layer_v3(mask, vertex, 64, rng="device")
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    for _ in range(3):
        rv.ransac_voting_layer_v3(mask, vertex, 64, rng="device")
    torch.cuda.synchronize()
    assert torch.cuda.memory_allocated() - before <= 1024          # only the [b,K,2] results
