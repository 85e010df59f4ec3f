#!/bin/bash
# round 2, GPU run 24: micro mixes 18/19 (software-pipelined); k_vote3 at HPL 4 x 3 CTAs with the sign-bit count
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -diag-suppress 128 -o gpurun_out/vote_mix benchmarks/micro/vote_mix.cu && timeout 300 gpurun_out/vote_mix > gpurun_out/micro_vote_mix.txt 2>&1
rm -f gpurun_out/vote_mix
rm -f gpurun_out/vote_sustained.jsonl gpurun_out/vote_sustained.err
for cfg in "4 3 4" "4 3 8" "8 2 4"; do
  set -- $cfg
  for field in planted random; do
    echo "# hpl $1 ctas $2 group $3" >> gpurun_out/vote_sustained.jsonl
    PVNET_VOTE_HPL=$1 PVNET_VOTE_CTAS=$2 PVNET_VOTE_GROUP=$3 SUST_FIELD=$field SUST_SKIP_BURST=1 timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
  done
done
echo finished > gpurun_out/run24.done
