#!/bin/bash
# round 2, GPU run 34b: half chunks (16-channel stages) for the resident 128-channel column layers (layer1, conv2s.0)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backbone.py -m gpu -q -rf --tb=short 2>&1 | tail -8 > gpurun_out/pytest_conv34.log
for half in 1 0; do
  PVNET_COL_HALF_CHUNKS=$half timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench34_half$half.json 2> gpurun_out/bench34_half$half.err
done
echo finished > gpurun_out/run34.done
