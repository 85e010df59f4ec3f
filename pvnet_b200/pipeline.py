"""End-to-end inference from host memory: image batches in pinned host buffers -> keypoints
(and optionally covariances) back on the host, the way tools/train_linemod.py:190-205
(`[d.cuda() for d in data]` -> net -> EvalWrapper -> `.cpu()`) runs the hot path, but with the
host->device copy of batch i+1 overlapped with the compute of batch i (two device input
buffers, a side stream for copies, CUDA events for ordering; no host synchronisation inside
the loop except the final one).
"""
from __future__ import annotations

import torch

from . import ransac_voting_gpu as rv


class PoseKeypointPipeline:
    def __init__(self, net, round_hyp_num=256, inlier_thresh=0.99, rng="batched", with_covariance=False,
                 cov_round_hyp_num=256, cov_min_hyp_num=4096):
        self.net = net
        self.hn = round_hyp_num
        self.thresh = inlier_thresh
        self.rng = rng
        self.with_cov = with_covariance
        self.cov_hn = cov_round_hyp_num
        self.cov_min = cov_min_hyp_num
        self._bufs = None
        self._copy_stream = None

    def _setup(self, host_batch, dev):
        if self._bufs is None or self._bufs[0].shape != host_batch.shape or self._bufs[0].device != dev:
            self._bufs = [torch.empty(host_batch.shape, dtype=torch.float32, device=dev) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]      # H2D of buffer i finished
            self._free = [torch.cuda.Event() for _ in range(2)]       # compute no longer reads buffer i
            self._copy_stream = torch.cuda.Stream(device=dev)
            for e in self._free:
                e.record(torch.cuda.current_stream(dev))

    def step(self, x):
        """x [b,3,H,W] on the device -> keypoints [b,K,2] (and covariances [b,K,2,2])."""
        out, mask = self.net.forward_native(x, with_mask=True)
        b, c, h, w = out.shape
        k = (c - self.net.seg_dim) // 2
        vertex = out[:, self.net.seg_dim:].permute(0, 2, 3, 1).view(b, h, w, k, 2)      # tools/demo.py:48-50
        kp = rv.ransac_voting_layer_v3(mask, vertex, self.hn, inlier_thresh=self.thresh, rng=self.rng)
        if not self.with_cov:
            return kp
        _, cov = rv.estimate_voting_distribution_with_mean(mask, vertex, kp, round_hyp_num=self.cov_hn,
                                                           min_hyp_num=self.cov_min, inlier_thresh=self.thresh,
                                                           rng=self.rng)
        return kp, cov

    @torch.no_grad()
    def run(self, host_batches, out_host=None, on_result=None):
        """host_batches: sequence of pinned float32 [b,3,H,W] tensors.  Results are copied
        device->host into out_host[i] (pinned) when given.  Returns the last device result."""
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
            result = self.step(self._bufs[j])
            self._free[j].record(main)
            if out_host is not None:
                kp = result[0] if isinstance(result, tuple) else result
                out_host[i].copy_(kp, non_blocking=True)
            if on_result is not None:
                on_result(i, result)
        return result
