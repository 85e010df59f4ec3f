"""GPU tests of the drop-in boundary under the reference's own wrappers and calling shapes:
nn.DataParallel over two devices (tools/train_linemod.py:258, tools/demo.py:160), copies / pickling
of a prepared module, raw uint8 input, CUDA-graph capture + replay of the whole hot path, and the
end-to-end PoseKeypointPipeline."""
import copy
import gc
import io

import numpy as np
import pytest
import torch
from torch import nn

from pvnet_b200 import ransac_voting_gpu as rv
from pvnet_b200 import synthetic as syn
from pvnet_b200.model_repository import Resnet18_8s
from pvnet_b200.pipeline import IMAGENET_MEAN, IMAGENET_STD, PoseKeypointPipeline
from tests.helpers import seeded_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(dev=DEV):
    net = Resnet18_8s(18, 2)
    net.load_state_dict(seeded_state_dict(net, 3))
    return net.to(dev).eval()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two visible GPUs")
def test_dataparallel_two_devices_shares_and_caches_handles():
    net = _net()
    x = torch.from_numpy(syn.backbone_input(4, 77, 96, 128)).to(DEV)
    with torch.no_grad():
        seg1, ver1 = net(x)
        assert net.native_pack_count() == 1
        dp = nn.DataParallel(net, device_ids=[0, 1])          # train_linemod.py:258
        seg, ver = dp(x)
        assert torch.equal(seg, seg1) and torch.equal(ver, ver1)
        assert net.native_pack_count() == 2                     # device 0 reused, device 1 packed once
        seg, ver = dp(x)
        assert net.native_pack_count() == 2, "weights were re-packed on the second DataParallel forward"
        assert torch.equal(seg, seg1)
        net.convraw[3].bias.add_(0.25)                           # in-place change: both devices re-pack once
        seg2, _ = dp(x)
        assert net.native_pack_count() == 4
        assert torch.allclose(seg2, seg1 + 0.25, atol=1e-5)
        del dp
        gc.collect()
        seg3, _ = net(x)                                         # the original's handle is still alive
        assert torch.equal(seg3, seg2)
    torch.cuda.synchronize()


def test_single_device_dataparallel_and_copies():
    net = _net()
    x = torch.from_numpy(syn.backbone_input(2, 5, 96, 128)).to(DEV)
    with torch.no_grad():
        seg, ver = net(x)
        shallow = copy.copy(net)                 # shares the native state: no double destroy, no re-pack
        s2, _ = shallow(x)
        assert torch.equal(s2, seg) and net.native_pack_count() == 1
        del shallow
        gc.collect()
        assert torch.equal(net(x)[0], seg)
        deep = copy.deepcopy(net)                # independent module, its own handle
        s3, _ = deep(x)
        assert torch.equal(s3, seg) and deep.native_pack_count() == 1
        buf = io.BytesIO()
        torch.save(net, buf)                     # a prepared module pickles (handles are dropped)
        buf.seek(0)
        loaded = torch.load(buf, weights_only=False)
        assert torch.equal(loaded(x)[0], seg)
        del deep, loaded
        gc.collect()
        assert torch.equal(net(x)[0], seg)
        frozen = net.freeze_native()
        assert torch.equal(frozen(x)[0], seg)
        net.freeze_native(False)


def test_uint8_input_is_bit_identical_to_torch_normalisation():
    net = _net()
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (2, 96, 128, 3), dtype=np.uint8)).to(DEV)
    mean = torch.tensor(IMAGENET_MEAN, device=DEV).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device=DEV).view(1, 3, 1, 1)
    xf = (img.permute(0, 3, 1, 2).float().div(255).sub(mean).div(std)).contiguous()     # ToTensor + Normalize
    with torch.no_grad():
        a, ma = net.forward_native(xf, with_mask=True, mask_dtype=torch.uint8)
        b, mb = net.forward_native(img, with_mask=True, mask_dtype=torch.uint8, mean=IMAGENET_MEAN, std=IMAGENET_STD)
    assert torch.equal(a, b) and torch.equal(ma, mb)
    with pytest.raises(ValueError):
        net.forward_native(img)


def test_cuda_graph_capture_and_replay_of_the_hot_path():
    """pvnet_backbone_forward + pvnet_ransac_voting_pipeline captured into one CUDA graph
    (include/pvnet_b200.h: nothing allocates or synchronises on the hot path)."""
    net = _net()
    x = torch.from_numpy(syn.backbone_input(2, 9)).to(DEV)

    def hot(xx):
        out, mask = net.forward_native(xx, with_mask=True, mask_dtype=torch.uint8)
        b, c, h, w = out.shape
        vertex = out[:, 2:].permute(0, 2, 3, 1).view(b, h, w, 9, 2)
        kp, cov = rv.ransac_voting_pipeline(mask, vertex, 64, 0.99, True, 64, 256, rng="device", min_num=0)
        return out, mask, kp, cov

    with torch.no_grad():
        torch.manual_seed(5)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):           # warm-up on the capture stream: plans, workspaces, attributes
            hot(x)
            rv.reset_device_rng(DEV)             # rewind the device RNG state in place
            e_out, e_mask, e_kp, e_cov = hot(x)  # eager result at offset 0
        side.synchronize()
        rv.reset_device_rng(DEV)                 # offset back to 0; same tensor address
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            g_out, g_mask, g_kp, g_cov = hot(x)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(g_out, e_out) and torch.equal(g_mask, e_mask)
        assert torch.equal(g_kp, e_kp) and torch.equal(g_cov, e_cov)      # same seed, same offset, same samples
        first = g_kp.clone()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(g_out, e_out)
        assert not torch.equal(g_kp, first), "a replay must draw fresh samples (device-side offset)"
        x.normal_()                               # new input through the same graph
        g.replay()
        torch.cuda.synchronize()
        assert not torch.equal(g_out, e_out) and torch.isfinite(g_out).all()


def test_pose_pipeline_run_from_host_buffers():
    net = _net()
    pipe = PoseKeypointPipeline(net, round_hyp_num=64, with_covariance=True, cov_round_hyp_num=64, cov_min_hyp_num=128)
    rng = np.random.default_rng(1)
    hosts = [torch.from_numpy(rng.integers(0, 256, (2, 96, 128, 3), dtype=np.uint8)).pin_memory() for _ in range(3)]
    kp_host = [torch.full([2, 9, 2], float("nan")).pin_memory() for _ in range(3)]
    cov_host = [torch.full([2, 9, 2, 2], float("nan")).pin_memory() for _ in range(3)]
    torch.manual_seed(11)
    rv.reset_device_rng(DEV)
    pipe.run(hosts, out_host=kp_host, cov_host=cov_host)
    # valid on return, without any synchronisation by the caller (run() used to return early)
    got_kp = [t.clone() for t in kp_host]
    got_cov = [t.clone() for t in cov_host]
    rv.reset_device_rng(DEV)
    with torch.no_grad():
        for i in range(3):
            kp, cov = pipe.step(hosts[i].to(DEV))
            assert torch.equal(kp.cpu(), got_kp[i]) and torch.equal(cov.cpu(), got_cov[i])
    # float32 NCHW host batches take the same path
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    hf = [(h.permute(0, 3, 1, 2).float().div(255).sub(mean).div(std)).contiguous().pin_memory() for h in hosts]
    rv.reset_device_rng(DEV)
    kp2 = [torch.empty([2, 9, 2]).pin_memory() for _ in range(3)]
    pipe.run(hf, out_host=kp2)
    for i in range(3):
        assert torch.equal(kp2[i], got_kp[i])


def test_pose_pipeline_graph_mode_equals_eager():
    """PoseKeypointPipeline(graph=True): each input buffer's device work is captured once and replayed; with the
    device sampler rewound the replays return exactly what the eager pipeline returns, and they keep drawing fresh
    samples when it is not rewound."""
    net = _net()
    rng = np.random.default_rng(2)
    pts3d = rng.uniform(-0.1, 0.1, (9, 3)).astype(np.float32)
    K = np.array([[572.4114, 0., 325.2611], [0., 573.57043, 242.04899], [0., 0., 1.]])
    kw = dict(round_hyp_num=64, with_covariance=True, cov_round_hyp_num=64, cov_min_hyp_num=128, points_3d=pts3d,
              camera_matrix=K)
    eager = PoseKeypointPipeline(net, **kw)
    graph = PoseKeypointPipeline(net, graph=True, **kw)
    hosts = [torch.from_numpy(rng.integers(0, 256, (2, 96, 128, 3), dtype=np.uint8)).pin_memory() for _ in range(5)]

    def run(pipe):
        kp = [torch.full([2, 9, 2], float("nan")).pin_memory() for _ in hosts]
        cov = [torch.full([2, 9, 2, 2], float("nan")).pin_memory() for _ in hosts]
        pose = [torch.full([2, 3, 4], float("nan"), dtype=torch.float64).pin_memory() for _ in hosts]
        rv.reset_device_rng(DEV)
        pipe.run(hosts, out_host=kp, cov_host=cov, pose_host=pose)
        return kp, cov, pose
    torch.manual_seed(13)
    run(graph)                                   # first run: warm-up + capture of the two graphs
    assert all(g is not None for g in graph._graphs)
    e = run(eager)
    g1 = run(graph)                              # pure replays
    for a, b in zip(e, g1):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    kp_a = [t.clone() for t in g1[0]]
    graph.run(hosts, out_host=g1[0])             # sampler not rewound: fresh samples
    assert any(not torch.equal(a, b) for a, b in zip(kp_a, g1[0]))
    with pytest.raises(ValueError):
        PoseKeypointPipeline(net, rng="batched", graph=True)


def test_pipeline_with_pose_and_pixel_major_equals_separate_calls():
    """PoseKeypointPipeline(points_3d, camera_matrix): keypoints / covariances / poses equal the ones obtained by
    calling the reference-shaped API step by step (NCHW output + permuted view, then v3 + with_mean with the same
    samples, then uncertainty_pnp) -- the pixel-major head output and the fused voting call change layouts, not values."""
    from pvnet_b200 import extend_utils as eu
    net = _net()
    rng = np.random.default_rng(4)
    pts3d = rng.uniform(-0.1, 0.1, (9, 3)).astype(np.float32)
    K = np.array([[572.4114, 0., 325.2611], [0., 573.57043, 242.04899], [0., 0., 1.]])
    pipe = PoseKeypointPipeline(net, round_hyp_num=64, with_covariance=True, cov_round_hyp_num=64, cov_min_hyp_num=128,
                                points_3d=pts3d, camera_matrix=K)
    x = torch.from_numpy(syn.backbone_input(2, 21, 96, 128)).to(DEV)
    with torch.no_grad():
        torch.manual_seed(3)
        rv.reset_device_rng(DEV)
        kp, cov, pose = pipe.step(x)
        out, mask = net.forward_native(x, with_mask=True)                       # NCHW, int64 mask (reference shapes)
        vertex = out[:, 2:].permute(0, 2, 3, 1).view(2, 96, 128, 9, 2)
        rv.reset_device_rng(DEV)
        kp2, cov2 = rv.ransac_voting_pipeline(mask, vertex, 64, 0.99, True, 64, 128, 0.99, rng="device")
        pose2 = eu.uncertainty_pnp_batched(kp2, pts3d, K, cov=cov2)
        pm = net.forward_native(x, pixel_major=True)
    assert torch.equal(pm.permute(0, 3, 1, 2), out)
    assert torch.equal(kp, kp2) and torch.equal(cov, cov2)
    assert pose.shape == (2, 3, 4) and pose.dtype == torch.float64
    assert torch.equal(torch.nan_to_num(pose), torch.nan_to_num(pose2))
