#!/bin/bash
# round 2, GPU run 14: k_vote3 (edge functionals, deferred guard band) against k_vote2; new reference-layer tests
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py tests/test_gpu_reference_layer.py tests/test_gpu_variants.py -m gpu -q -rf --tb=short 2>&1 | tail -30 > gpurun_out/pytest_vote3.log
for cfg in "2 8 2 8" "3 8 2 4" "3 8 2 8" "3 4 3 4" "3 4 3 8"; do
  set -- $cfg
  PVNET_VOTE_IMPL=$1 PVNET_VOTE_HPL=$2 PVNET_VOTE_CTAS=$3 PVNET_VOTE_GROUP=$4 SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 \
    python benchmarks/vote_sweep.py > gpurun_out/sweep14_i$1_hpl$2_c$3_g$4.jsonl 2> gpurun_out/sweep14_i$1_hpl$2_c$3_g$4.err
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench14_n1.json 2> gpurun_out/bench14_n1.err
PVNET_VOTE_IMPL=2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench14_n1_impl2.json 2> gpurun_out/bench14_n1_impl2.err
echo finished > gpurun_out/run14.done
