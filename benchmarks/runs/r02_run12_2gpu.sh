#!/bin/bash
# round 2, GPU run 12 (2 GPUs): DataParallel test again, nvJPEG tests
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_jpeg.py -m gpu -q -rf --tb=short 2>&1 | tail -30 > gpurun_out/pytest_2gpu.log
echo finished > gpurun_out/run12.done
