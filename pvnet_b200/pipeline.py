"""End-to-end inference from host memory: image batches in pinned host buffers -> keypoints
(and optionally covariances) back on the host, the way tools/train_linemod.py:190-205
(`[d.cuda() for d in data]` -> net -> EvalWrapper / UncertaintyEvalWrapper -> `.cpu()`) runs the
hot path, but with the host->device copy of batch i+1 overlapped with the compute of batch i
(two device input buffers, a side stream for copies, CUDA events for ordering; no host
synchronisation inside the loop; `run` synchronises once on the last device->host copy before it
returns, so the host buffers are valid when it does).

Per batch the device work is: `pvnet_backbone_forward` (fused argmax -> uint8 mask) and ONE
`pvnet_ransac_voting_pipeline` call (v3, plus estimate_voting_distribution_with_mean when
`with_covariance`), sampling on the device (rng="device": no torch RNG launches).

With `points_3d` + `camera_matrix` the uncertainty-driven PnP of `Evaluator.evaluate_uncertainty`
(lib/utils/evaluation_utils.py:165-201) runs on the device as well (`pvnet_uncertainty_pnp`), so POSES
[b,3,4] are what leaves the GPU (and what an 8-GPU job gathers).

`graph=True`: the per-batch device work (31 backbone launches + the voting call's ~10 + PnP) is captured into one
CUDA graph per input buffer on first use and replayed afterwards -- 4 us of host time per batch instead of
0.4-5 ms of Python + launches (batch 1: 0.72 ms per image instead of 0.78; at batch 16 the GPU is the limit either
way, the host thread is what is freed).  The device-side sampler keeps drawing fresh samples across replays.

Inputs may be float32 [b,3,H,W] (already normalised, what `ToTensor` + `Normalize` produce,
tools/demo.py:89-95) or uint8 [b,H,W,3] raw images: the latter are normalised on the device inside
the packing kernel (4x fewer host->device bytes).
"""
from __future__ import annotations

import torch

from . import extend_utils as eu
from . import ransac_voting_gpu as rv

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # tools/demo.py:91-94, lib/datasets/linemod_dataset.py:191-195
IMAGENET_STD = (0.229, 0.224, 0.225)


class PoseKeypointPipeline:
    def __init__(self, net, round_hyp_num=256, inlier_thresh=0.99, rng="device", with_covariance=False,
                 cov_round_hyp_num=256, cov_min_hyp_num=4096, max_num=30000, mean=IMAGENET_MEAN, std=IMAGENET_STD,
                 points_3d=None, camera_matrix=None, graph=False):
        self.net = net
        self.graph = bool(graph)
        self.hn = round_hyp_num
        self.thresh = inlier_thresh
        self.rng = rng
        self.with_cov = with_covariance
        self.cov_hn = cov_round_hyp_num
        self.cov_min = cov_min_hyp_num
        self.max_num = max_num
        self.mean, self.std = tuple(mean), tuple(std)
        self.with_pose = points_3d is not None and camera_matrix is not None
        if self.with_pose and not with_covariance:
            raise ValueError("poses need the covariances: with_covariance=True")
        if self.graph and rng != "device":
            raise ValueError("graph=True needs rng='device' (torch's generator cannot be replayed)")
        self.points_3d, self.camera_matrix = points_3d, camera_matrix
        self._p3_dev = None
        self._bufs = None
        self._copy_stream = None

    def _setup(self, host_batch, dev):
        if (self._bufs is None or self._bufs[0].shape != host_batch.shape or self._bufs[0].dtype != host_batch.dtype
                or self._bufs[0].device != dev):
            self._bufs = [torch.empty(host_batch.shape, dtype=host_batch.dtype, device=dev) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]      # H2D of buffer i finished
            self._free = [torch.cuda.Event() for _ in range(2)]       # compute no longer reads buffer i
            self._done = torch.cuda.Event()                           # last D2H of a run() finished
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._graphs = [None, None]                               # per input buffer: (CUDAGraph, static result)
            for e in self._free:
                e.record(torch.cuda.current_stream(dev))

    def _step_graph(self, j):
        """Replay (capture on first use) the graph of `step(self._bufs[j])` on the current stream."""
        if self._graphs[j] is None:
            cur = torch.cuda.current_stream(self._bufs[j].device)
            side = torch.cuda.Stream(device=self._bufs[j].device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.step(self._bufs[j])            # eager once on the capture stream: plans, workspaces, attributes
                side.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    res = self.step(self._bufs[j])
            cur.wait_stream(side)
            self._graphs[j] = (g, res)
        g, res = self._graphs[j]
        g.replay()
        return res

    def step(self, x):
        """x on the device: float32 [b,3,H,W] or uint8 [b,H,W,3] -> keypoints [b,K,2]
        (and covariances [b,K,2,2]) (and poses [b,3,4] float64)."""
        # pixel-major head output: the vertex field is the contiguous [b,h,w,K,2] form of the permuted view of
        # tools/demo.py:48-50 (same values; the voting layer's gather then reads whole records, not sectors)
        if x.dtype == torch.uint8:
            out, mask = self.net.forward_native(x, with_mask=True, mask_dtype=torch.uint8, mean=self.mean, std=self.std,
                                                pixel_major=True)
        else:
            out, mask = self.net.forward_native(x, with_mask=True, mask_dtype=torch.uint8, pixel_major=True)
        b, h, w, c = out.shape
        k = (c - self.net.seg_dim) // 2
        vertex = out[..., self.net.seg_dim:].unflatten(3, (k, 2))
        # a 2-class argmax mask is binary: v3's `nonzero` and with_mean's `== 1` readings coincide
        if self.net.seg_dim == 2 or not self.with_cov:
            res = rv.ransac_voting_pipeline(mask, vertex, self.hn, self.thresh, self.with_cov, self.cov_hn, self.cov_min,
                                            self.thresh, max_num=self.max_num, rng=self.rng)
        else:
            kp = rv.ransac_voting_layer_v3(mask, vertex, self.hn, inlier_thresh=self.thresh, max_num=self.max_num,
                                           rng="batched")
            _, cov = rv.estimate_voting_distribution_with_mean(mask, vertex, kp, round_hyp_num=self.cov_hn,
                                                               min_hyp_num=self.cov_min, inlier_thresh=self.thresh,
                                                               max_num=self.max_num, rng="batched")
            res = (kp, cov)
        if self.with_pose:
            if self._p3_dev is None or self._p3_dev.device != x.device:     # the model points go to the device once
                self._p3_dev = torch.as_tensor(self.points_3d, dtype=torch.float32).to(x.device).contiguous()
            pose = eu.uncertainty_pnp_batched(res[0], self._p3_dev, self.camera_matrix, cov=res[1])
            return res[0], res[1], pose
        return res

    @torch.no_grad()
    def run(self, host_batches, out_host=None, cov_host=None, on_result=None, pose_host=None):
        """host_batches: sequence of pinned [b,3,H,W] float32 (or [b,H,W,3] uint8) tensors.  Results are
        copied device->host into out_host[i] (and cov_host[i]) -- pinned tensors -- when given; the call
        returns after the last of those copies has completed.  Returns the last device result (with graph=True a
        static tensor that the next replay on the same input buffer overwrites)."""
        dev = next(self.net.parameters()).device
        batches = list(host_batches)
        if not batches:
            return None
        self._setup(batches[0], dev)
        main = torch.cuda.current_stream(dev)
        cs = self._copy_stream
        result = None

        def upload(i):
            j = i & 1
            cs.wait_event(self._free[j])
            with torch.cuda.stream(cs):
                self._bufs[j].copy_(batches[i], non_blocking=True)
                self._ready[j].record(cs)
        upload(0)
        for i in range(len(batches)):
            j = i & 1
            if i + 1 < len(batches):
                upload(i + 1)
            main.wait_event(self._ready[j])
            result = self._step_graph(j) if self.graph else self.step(self._bufs[j])
            self._free[j].record(main)
            if out_host is not None:
                kp = result[0] if isinstance(result, tuple) else result
                out_host[i].copy_(kp, non_blocking=True)
            if cov_host is not None and isinstance(result, tuple):
                cov_host[i].copy_(result[1], non_blocking=True)
            if pose_host is not None and isinstance(result, tuple) and len(result) > 2:
                pose_host[i].copy_(result[2], non_blocking=True)
            if on_result is not None:
                on_result(i, result)
        if out_host is not None or cov_host is not None or pose_host is not None:
            self._done.record(main)
            self._done.synchronize()
        return result
