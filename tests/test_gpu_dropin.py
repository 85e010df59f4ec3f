"""GPU: the drop-in acceptance path.  tools/demo.py and tools/train_linemod.py cannot be imported in
this image (matplotlib, easydict, transforms3d, plyfile, tensorboardX are absent), so these tests run
the exact call sequences of their wrappers -- EvalWrapper (train_linemod.py:94-106), its
use_uncertainty branch (:104), UncertaintyEvalWrapper (:119-130), demo.py's generate_hypothesis call
(:121-125) -- through the reference's import paths (`lib.*`), with DataParallel as the reference
uses it (train_linemod.py:183-184)."""
import numpy as np
import pytest
import torch
from torch import nn

from lib.networks.model_repository import *          # noqa: F401,F403  (the reference's import line)
from lib.ransac_voting_gpu_layer.ransac_voting_gpu import (estimate_voting_distribution_with_mean, generate_hypothesis,
                                                           ransac_voting_layer_v3, ransac_voting_layer_v5)
from pvnet_b200 import synthetic as syn
from tests.helpers import demo_fixture, seeded_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class EvalWrapper(nn.Module):
    def forward(self, seg_pred, vertex_pred, use_argmax=True, use_uncertainty=False):
        vertex_pred = vertex_pred.permute(0, 2, 3, 1)
        b, h, w, vn_2 = vertex_pred.shape
        vertex_pred = vertex_pred.view(b, h, w, vn_2 // 2, 2)
        mask = torch.argmax(seg_pred, 1) if use_argmax else seg_pred
        if use_uncertainty:
            return ransac_voting_layer_v5(mask, vertex_pred, 128, inlier_thresh=0.99, max_num=100)
        return ransac_voting_layer_v3(mask, vertex_pred, 128, inlier_thresh=0.99, max_num=100)


class UncertaintyEvalWrapper(nn.Module):
    def forward(self, seg_pred, vertex_pred, use_argmax=True):
        vertex_pred = vertex_pred.permute(0, 2, 3, 1)
        b, h, w, vn_2 = vertex_pred.shape
        vertex_pred = vertex_pred.view(b, h, w, vn_2 // 2, 2)
        mask = torch.argmax(seg_pred, 1) if use_argmax else seg_pred
        mean = ransac_voting_layer_v3(mask, vertex_pred, 512, inlier_thresh=0.99)
        mean, var = estimate_voting_distribution_with_mean(mask, vertex_pred, mean)
        return mean, var


def _demo_batch():
    """seg logits / vertex field a perfect network would emit for the reference's demo image."""
    mask, field, pts = demo_fixture()
    seg = torch.from_numpy(np.stack([1.0 - mask, mask.astype(np.float64)]).astype(np.float32))[None]
    ver = torch.from_numpy(field)[None]
    return seg.to(DEV), ver.to(DEV), pts


def test_eval_wrapper_sequences_on_demo_fixture():
    seg, ver, pts = _demo_batch()
    torch.manual_seed(0)
    eval_net = nn.DataParallel(EvalWrapper().to(DEV), device_ids=[0])
    kp = eval_net(seg, ver).cpu().numpy()
    assert kp.shape == (1, 9, 2) and np.abs(kp[0] - pts).max() < 0.1      # max_num=100 subsamples to ~100 px
    kp5, conf = eval_net(seg, ver, True, True)
    assert kp5.shape == (1, 9, 2) and conf.shape == (1, 9) and float(conf.min()) > 0.5
    unc_net = nn.DataParallel(UncertaintyEvalWrapper().to(DEV), device_ids=[0])
    mean, var = unc_net(seg, ver)
    assert np.abs(mean.cpu().numpy()[0] - pts).max() < 1e-2 and var.shape == (1, 9, 2, 2)
    assert torch.isfinite(var).all()


def test_demo_generate_hypothesis_call():
    seg, ver, _ = _demo_batch()
    vertex = ver.permute(0, 2, 3, 1)
    b, h, w, vn_2 = vertex.shape
    vertex = vertex.view(b, h, w, vn_2 // 2, 2)
    mask = torch.argmax(seg, 1)
    hyp, cnt = generate_hypothesis(mask, vertex, 128, inlier_thresh=0.99)      # demo.py:121-125
    assert hyp.shape == (1, 128, 9, 2) and cnt.shape == (1, 128, 9) and cnt.dtype == torch.int64
    assert int(cnt.max()) > 2000


def test_full_pipeline_through_reference_import_paths():
    net = Resnet18_8s(ver_dim=18, seg_dim=2)          # noqa: F405 (star import, as tools/demo.py:158)
    net.load_state_dict(seeded_state_dict(net, seed=1))
    net = net.to(DEV).eval()
    x = torch.from_numpy(syn.backbone_input(2, 4, 240, 320)).to(DEV)
    with torch.no_grad():
        seg_pred, vertex_pred = net(x)
        kp = EvalWrapper()(seg_pred, vertex_pred)
    assert kp.shape == (2, 9, 2) and torch.isfinite(kp).all()
