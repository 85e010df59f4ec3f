#!/bin/bash
# round 2, GPU run 36b: ncu --set full of the 25 tensor-core conv launches of one step (conv table with the tc-pipe columns)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 800 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_conv_tap_p|k_conv_col|k_conv_tc" -c 25 \
  -o gpurun_out/conv_full -f python benchmarks/profile_step.py 1 > gpurun_out/ncu_conv_full.log 2>&1
echo finished > gpurun_out/run36.done
