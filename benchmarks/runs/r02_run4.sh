#!/bin/bash
# round 2, GPU run 4: instruction-mix microbenchmark, GPU-vs-GPU baselines, b=1 latency, new tests
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/vote_mix benchmarks/micro/vote_mix.cu 2> gpurun_out/vote_mix_build.log
timeout 120 ./gpurun_out/vote_mix > gpurun_out/micro_vote_mix.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_pipeline.py tests/test_gpu_pnp.py tests/test_gpu_boundary.py -m gpu -q -rf --tb=short 2>&1 | tail -60 > gpurun_out/pytest_new.log
timeout 600 python benchmarks/gpu_baselines.py > gpurun_out/gpu_baselines.jsonl 2> gpurun_out/gpu_baselines.err
timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b1.jsonl 2> gpurun_out/latency_b1.err
LAT_BATCH=16 timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b16.jsonl 2> gpurun_out/latency_b16.err
echo finished > gpurun_out/run4.done
