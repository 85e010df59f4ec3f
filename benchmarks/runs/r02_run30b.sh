#!/bin/bash
# round 2, GPU run 30b: ncu --set full of convraw.0, fused-upsampling variant and separate variant
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_col --launch-skip 7 -c 1 \
  -o gpurun_out/convraw_fused -f python benchmarks/profile_step.py 1 > gpurun_out/ncu_convraw_fused.log 2>&1
PVNET_FUSE_UP=0 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_col --launch-skip 7 -c 1 \
  -o gpurun_out/convraw_separate -f python benchmarks/profile_step.py 1 > gpurun_out/ncu_convraw_separate.log 2>&1
echo finished > gpurun_out/run30.done
