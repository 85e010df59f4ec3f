#!/bin/bash
# round 2, GPU run 5: full tests, occupancy A/B of k_vote2, bench, baselines, latency
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
for cfg in "4 4" "4 3" "8 2"; do
  set -- $cfg
  PVNET_VOTE_HPL=$1 PVNET_VOTE_CTAS=$2 SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 \
    python benchmarks/vote_sweep.py > gpurun_out/sweep5_hpl$1_c$2.jsonl 2> gpurun_out/sweep5_hpl$1_c$2.err
done
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench5_n1.json 2> gpurun_out/bench5_n1.err
PVNET_VOTE_HPL=4 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench5_n1_hpl4.json 2> gpurun_out/bench5_n1_hpl4.err
timeout 600 python benchmarks/gpu_baselines.py > gpurun_out/gpu_baselines.jsonl 2> gpurun_out/gpu_baselines.err
timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b1.jsonl 2> gpurun_out/latency_b1.err
echo finished > gpurun_out/run5.done
