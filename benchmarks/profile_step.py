"""A bare loop of the bench step (backbone + vote, batch 16) for ncu: no calibration, no
per-stage timing, no CPU baseline.  `python benchmarks/profile_step.py [steps]`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvnet_b200 import synthetic as syn  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda", 0)
    net = bench.build_model(torch, dev)
    with torch.no_grad():
        net.convraw[3].bias[1] -= 1.95       # ~6.5% foreground without the calibration pass
    step = bench.make_step(torch, net)
    x = torch.from_numpy(syn.backbone_input(bench.BATCH, 2000)).to(dev)
    with torch.no_grad():
        for _ in range(steps):
            step(x)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
