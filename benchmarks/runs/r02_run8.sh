#!/bin/bash
# round 2, GPU run 8: per-group check restored (G = 4 / 8), PnP on well-posed inputs, bench, ncu
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py tests/test_gpu_reference_layer.py -m gpu -q -rf --tb=short 2>&1 | tail -30 > gpurun_out/pytest_vote.log
for cfg in "8 2 4" "8 2 8" "4 3 4" "4 3 8"; do
  set -- $cfg
  PVNET_VOTE_HPL=$1 PVNET_VOTE_CTAS=$2 PVNET_VOTE_GROUP=$3 SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 \
    python benchmarks/vote_sweep.py > gpurun_out/sweep8_hpl$1_c$2_g$3.jsonl 2> gpurun_out/sweep8_hpl$1_c$2_g$3.err
done
LAT_ONLY_PNP=1 timeout 200 python benchmarks/latency_b1.py > gpurun_out/pnp_timing.jsonl 2> gpurun_out/pnp_timing.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench8_n1.json 2> gpurun_out/bench8_n1.err
PVNET_VOTE_GROUP=8 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench8_n1_g8.json 2> gpurun_out/bench8_n1_g8.err
echo finished > gpurun_out/run8.done
