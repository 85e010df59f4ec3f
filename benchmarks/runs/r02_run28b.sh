#!/bin/bash
# round 2, GPU run 28b (and 29b: loads batched, 31b: lean interpolation): 1/2 -> 1 upsampling fused into convraw.0's loader -- equality with the separate launch, backbone tests, bench A/B
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backbone.py -m gpu -q -rf --tb=short -s -k fused_upsample 2>&1 | tail -40 > gpurun_out/pytest_fused_up.log
if grep -q "passed" gpurun_out/pytest_fused_up.log && ! grep -q "failed" gpurun_out/pytest_fused_up.log; then
  timeout 600 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_boundary.py tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -m gpu -q -rf --tb=short 2>&1 | tail -15 > gpurun_out/pytest_backbone28.log
fi
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench28_fused.json 2> gpurun_out/bench28_fused.err
PVNET_FUSE_UP=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench28_separate.json 2> gpurun_out/bench28_separate.err
echo finished > gpurun_out/run28.done
