"""Batch-1 latency of the hot path (the reference's own calling shape: tools/demo.py and
train_linemod.py's test_batch_size = 1): Resnet18_8s forward + fused argmax + v3(256 hyp)
[+ with_mean(256, 4096)] [+ uncertainty PnP], eager launches vs one CUDA graph replay.
Device-resident input; per-iteration time = CUDA events around 200 back-to-back iterations; also the
host-side enqueue time of one eager step.  Prints one JSON line per variant."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvnet_b200 import extend_utils as eu  # noqa: E402
from pvnet_b200 import ransac_voting_gpu as rv  # noqa: E402
from pvnet_b200 import synthetic as syn  # noqa: E402

DEV = torch.device("cuda", 0)
K_MAT = np.array([[572.4114, 0., 325.2611], [0., 573.57043, 242.04899], [0., 0., 1.]])


def main():
    b = int(os.environ.get("LAT_BATCH", "1"))
    net = bench.build_model(torch, DEV).freeze_native(False)
    x = torch.from_numpy(syn.backbone_input(max(b, 4), 3)).to(DEV)
    bench.calibrate_foreground(torch, net, x)
    x = x[:b].contiguous()
    pts3d = torch.from_numpy(np.random.default_rng(0).uniform(-0.1, 0.1, (9, 3)).astype(np.float32)).to(DEV)

    def make(with_cov, with_pose):
        def step():
            out, mask = net.forward_native(x, with_mask=True, mask_dtype=torch.uint8)
            vertex = out[:, 2:].permute(0, 2, 3, 1).view(b, 480, 640, 9, 2)
            r = rv.ransac_voting_pipeline(mask, vertex, 256, 0.99, with_cov, 256, 4096, 0.99, rng="device")
            if with_pose:
                return eu.uncertainty_pnp_batched(r[0], pts3d, K_MAT, cov=r[1])
            return r
        return step

    for name, with_cov, with_pose in (("backbone+v3", False, False), ("backbone+v3+cov", True, False),
                                      ("backbone+v3+cov+pnp", True, True)):
        step = make(with_cov, with_pose)
        with torch.no_grad():
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            side.synchronize()
            net.freeze_native(True)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                step()
            row = {"variant": name, "batch": b}
            for mode, fn in (("eager", step), ("graph", g.replay)):
                with torch.cuda.stream(side):
                    for _ in range(10):
                        fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    n = 200
                    t0 = time.perf_counter()
                    e0.record()
                    for _ in range(n):
                        fn()
                    e1.record()
                    t_enq = (time.perf_counter() - t0) / n * 1e3
                side.synchronize()
                row[f"{mode}_ms_per_iter"] = round(e0.elapsed_time(e1) / n, 4)
                row[f"{mode}_host_enqueue_ms"] = round(t_enq, 4)
            net.freeze_native(False)
            row["images_per_s_graph"] = round(b / row["graph_ms_per_iter"] * 1e3, 1)
        print(json.dumps(row), flush=True)


def pnp_timing():
    """The device PnP alone on well-posed inputs (the network above has random weights: its keypoints are
    noise, so the LM there runs into its iteration cap): object keypoints under random poses + anisotropic
    noise, batch 1 / 16 / 128; CUDA events over 50 calls; mean LM iterations from the info output."""
    rng = np.random.default_rng(0)
    pts = rng.uniform(-0.1, 0.1, (9, 3)).astype(np.float32)

    def rod(r):
        th = np.linalg.norm(r)
        k = r / th
        kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(th) * kx + (1 - np.cos(th)) * kx @ kx
    for b in (1, 16, 128):
        kps, covs = [], []
        for _ in range(b):
            R, t = rod(rng.normal(0, 1, 3)), np.array([rng.uniform(-.1, .1), rng.uniform(-.1, .1), rng.uniform(.6, 1.2)])
            X = pts @ R.T + t
            uv = np.stack([K_MAT[0, 0] * X[:, 0] / X[:, 2] + K_MAT[0, 2], K_MAT[1, 1] * X[:, 1] / X[:, 2] + K_MAT[1, 2]], 1)
            A = rng.normal(0, 1, (9, 2, 2))
            cov = A @ A.transpose(0, 2, 1) + 0.3 * np.eye(2)
            kps.append(uv + np.stack([rng.multivariate_normal(np.zeros(2), c) for c in cov]))
            covs.append(cov)
        kp = torch.from_numpy(np.stack(kps).astype(np.float32)).to(DEV)
        cov = torch.from_numpy(np.stack(covs).astype(np.float32)).to(DEV)
        p3 = torch.from_numpy(pts).to(DEV)
        for _ in range(5):
            _, info = eu.uncertainty_pnp_batched(kp, p3, K_MAT, cov=cov, return_info=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            eu.uncertainty_pnp_batched(kp, p3, K_MAT, cov=cov)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"variant": "uncertainty_pnp alone, well-posed inputs", "batch": b,
                          "ms_per_call": round(e0.elapsed_time(e1) / 50, 4),
                          "lm_iterations_mean": round(float(info[:, 1].float().mean()), 2),
                          "status_nonzero": int((info[:, 0] != 0).sum())}), flush=True)


if __name__ == "__main__":
    if os.environ.get("LAT_ONLY_PNP") != "1":
        main()
    pnp_timing()
