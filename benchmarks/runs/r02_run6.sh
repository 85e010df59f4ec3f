#!/bin/bash
# round 2, GPU run 6: branch-free sweep; overlap experiment; PnP timing
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py tests/test_gpu_pnp.py tests/test_gpu_reference_layer.py tests/test_gpu_variants.py -m gpu -q -rf --tb=short 2>&1 | tail -40 > gpurun_out/pytest_vote.log
for cfg in "4 4" "8 2"; do
  set -- $cfg
  PVNET_VOTE_HPL=$1 PVNET_VOTE_CTAS=$2 SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 \
    python benchmarks/vote_sweep.py > gpurun_out/sweep6_hpl$1_c$2.jsonl 2> gpurun_out/sweep6_hpl$1_c$2.err
done
timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b1.jsonl 2> gpurun_out/latency_b1.err
: > gpurun_out/overlap.jsonl
timeout 200 python benchmarks/overlap_experiment.py >> gpurun_out/overlap.jsonl 2> gpurun_out/overlap.err
PVNET_CONV_STAGES=3 timeout 200 python benchmarks/overlap_experiment.py >> gpurun_out/overlap.jsonl 2>> gpurun_out/overlap.err
PVNET_CONV_STAGES=3 PVNET_VOTE_HPL=8 PVNET_VOTE_CTAS=1 timeout 200 python benchmarks/overlap_experiment.py >> gpurun_out/overlap.jsonl 2>> gpurun_out/overlap.err
PVNET_CONV_STAGES=3 PVNET_VOTE_HPL=4 PVNET_VOTE_CTAS=2 timeout 200 python benchmarks/overlap_experiment.py >> gpurun_out/overlap.jsonl 2>> gpurun_out/overlap.err
PVNET_CONV_STAGES=3 OVL_COV=0 timeout 200 python benchmarks/overlap_experiment.py >> gpurun_out/overlap.jsonl 2>> gpurun_out/overlap.err
echo finished > gpurun_out/run6.done
