#!/bin/bash
# round 2, GPU run 23: ncu --set full of k_vote3 at the config-4 shape (planted field)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
SUST_FIELD=planted SUST_SKIP_BURST=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vote3 --launch-skip 6 -c 1 \
  -o gpurun_out/vote3_cfg4 -f python benchmarks/vote_sustained.py > gpurun_out/ncu_vote3_cfg4.log 2>&1
echo finished > gpurun_out/run23.done
