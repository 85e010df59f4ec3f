"""Device time of the bench step split into backbone and voting, measured in place with CUDA
events recorded between the two calls of every step (30 steps after 10 warm-up steps), plus the
CPU time spent enqueueing a step.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvnet_b200 import ransac_voting_gpu as rv  # noqa: E402
from pvnet_b200 import synthetic as syn  # noqa: E402


WITH_COV = os.environ.get("BREAKDOWN_COV", "0") == "1"


def main():
    dev = torch.device("cuda", 0)
    net = bench.build_model(torch, dev)
    xs = [torch.from_numpy(syn.backbone_input(bench.BATCH, 2000 + i)).to(dev) for i in range(3)]
    bench.calibrate_foreground(torch, net, xs[0])
    steps, warm = 30, 10
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    cpu = []
    with torch.no_grad():
        for i in range(warm + steps):
            x = xs[i % 3]
            t0 = time.perf_counter()
            if i >= warm:
                ev[i - warm][0].record()
            out, mask = net.forward_native(x, with_mask=True, mask_dtype=torch.uint8, pixel_major=True)
            if i >= warm:
                ev[i - warm][1].record()
            vertex = out[..., 2:].unflatten(3, (bench.K_KP, 2))
            rv.ransac_voting_pipeline(mask, vertex, bench.HYP, bench.THRESH, WITH_COV, 256, 4096, bench.THRESH, rng="device")
            if i >= warm:
                ev[i - warm][2].record()
            cpu.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
    bb = [e[0].elapsed_time(e[1]) for e in ev]
    vt = [e[1].elapsed_time(e[2]) for e in ev]
    total = ev[0][0].elapsed_time(ev[-1][2]) / steps
    print(json.dumps(dict(with_cov=WITH_COV, backbone_ms=round(float(np.median(bb)), 4), vote_ms=round(float(np.median(vt)), 4),
                          step_ms=round(total, 4), cpu_enqueue_ms=round(float(np.median(cpu[warm:])) * 1e3, 4))))


if __name__ == "__main__":
    main()
