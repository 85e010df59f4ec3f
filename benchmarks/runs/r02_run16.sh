#!/bin/bash
# round 2, GPU run 16 (and 21): micro mixes
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -diag-suppress 128 -o gpurun_out/vote_mix benchmarks/micro/vote_mix.cu && timeout 200 gpurun_out/vote_mix > gpurun_out/micro_vote_mix.txt 2>&1
rm -f gpurun_out/vote_mix
echo finished > gpurun_out/run16.done
