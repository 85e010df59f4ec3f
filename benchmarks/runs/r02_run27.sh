#!/bin/bash
# round 2, GPU run 27: GPU-vs-GPU baselines again (runs 20/26 had them under an ncu wrapper by a script slip: discarded)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 600 python benchmarks/gpu_baselines.py > gpurun_out/gpu_baselines.jsonl 2> gpurun_out/gpu_baselines.err
echo finished > gpurun_out/run27.done
