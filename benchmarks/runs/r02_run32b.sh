#!/bin/bash
# round 2, GPU run 32b: convraw.0 with two epilogue warp sets (PVNET_HEAD_EPI) x fused / separate upsampling (PVNET_FUSE_UP)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_boundary.py -m gpu -q -rf --tb=short 2>&1 | tail -8 > gpurun_out/pytest_backbone32.log
for cfg in "1 2" "0 2" "0 1" "1 1"; do
  set -- $cfg
  PVNET_FUSE_UP=$1 PVNET_HEAD_EPI=$2 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench32_up$1_epi$2.json 2> gpurun_out/bench32_up$1_epi$2.err
done
echo finished > gpurun_out/run32.done
