#!/bin/bash
# round 2, GPU run 35: PoseKeypointPipeline(graph=True)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py -m gpu -q -rf --tb=short 2>&1 | tail -25 > gpurun_out/pytest_graphpipe.log
echo finished > gpurun_out/run35.done
