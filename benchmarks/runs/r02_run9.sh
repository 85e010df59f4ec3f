#!/bin/bash
# round 2, GPU run 9: common per-lane band + FMNMX3 band check; occupancy 2 vs 3; small-batch N tiles; b=1 latency
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
for cfg in "8 2 8" "8 3 8" "8 3 4" "4 3 8"; do
  set -- $cfg
  PVNET_VOTE_HPL=$1 PVNET_VOTE_CTAS=$2 PVNET_VOTE_GROUP=$3 SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 \
    python benchmarks/vote_sweep.py > gpurun_out/sweep9_hpl$1_c$2_g$3.jsonl 2> gpurun_out/sweep9_hpl$1_c$2_g$3.err
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench9_n1.json 2> gpurun_out/bench9_n1.err
PVNET_VOTE_CTAS=3 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench9_n1_c3.json 2> gpurun_out/bench9_n1_c3.err
timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b1.jsonl 2> gpurun_out/latency_b1.err
PVNET_CONV_SMALL_BATCH_SPLIT=0 LAT_ONLY=1 timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b1_nosplit.jsonl 2> gpurun_out/latency_b1_nosplit.err
echo finished > gpurun_out/run9.done
