#!/bin/bash
# round 2, GPU run 15: is the config-4 voting layer power-capped? (k_vote2 vs k_vote3, burst vs sustained); new micro mixes
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -diag-suppress 128 -o gpurun_out/vote_mix benchmarks/micro/vote_mix.cu && timeout 120 gpurun_out/vote_mix > gpurun_out/micro_vote_mix.txt 2>&1
rm -f gpurun_out/vote_mix
for impl in 2 3; do
  for field in planted random; do
    PVNET_VOTE_IMPL=$impl SUST_FIELD=$field timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
  done
done
PVNET_VOTE_IMPL=3 PVNET_VOTE_GROUP=8 SUST_FIELD=planted timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
echo finished > gpurun_out/run15.done
