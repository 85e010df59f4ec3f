#!/bin/bash
# round 2, GPU run 38b (and 40b, with the L2 prefetch in the interpolation warps): committed default path once more (full gpu tests, smoke, bench), then the fused upsampling with eight
# dedicated interpolation warps (PVNET_FUSE_UP=2): bit-exactness tests, bench A/B
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rf --tb=short -k "not dedicated" 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench38_default.json 2> gpurun_out/bench38_default.err
timeout 150 python -m pytest tests/test_gpu_backbone.py -m gpu -q -rf --tb=short -s -k "dedicated" 2>&1 | tail -30 > gpurun_out/pytest_dedicated.log
if grep -q " passed" gpurun_out/pytest_dedicated.log && ! grep -q "failed\|error" gpurun_out/pytest_dedicated.log; then
  PVNET_FUSE_UP=2 timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench38_up2.json 2> gpurun_out/bench38_up2.err
fi
echo finished > gpurun_out/run38.done
