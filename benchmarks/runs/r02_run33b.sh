#!/bin/bash
# round 2, GPU run 33b: ncu --set full of the fused-upsampling convraw.0 (lean interpolation, one epilogue set)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
PVNET_FUSE_UP=1 PVNET_HEAD_EPI=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_col --launch-skip 7 -c 1 \
  -o gpurun_out/convraw_fused_v3 -f python benchmarks/profile_step.py 1 > gpurun_out/ncu_convraw_fused_v3.log 2>&1
echo finished > gpurun_out/run33.done
