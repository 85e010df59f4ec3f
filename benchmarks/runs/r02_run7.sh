#!/bin/bash
# round 2, GPU run 7: per-pixel band bits in k_vote2; PnP timing
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py tests/test_gpu_pnp.py tests/test_gpu_reference_layer.py tests/test_gpu_variants.py -m gpu -q -rf --tb=short 2>&1 | tail -40 > gpurun_out/pytest_vote.log
for cfg in "4 3" "8 2"; do
  set -- $cfg
  PVNET_VOTE_HPL=$1 PVNET_VOTE_CTAS=$2 SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 \
    python benchmarks/vote_sweep.py > gpurun_out/sweep7_hpl$1_c$2.jsonl 2> gpurun_out/sweep7_hpl$1_c$2.err
done
timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b1.jsonl 2> gpurun_out/latency_b1.err
LAT_BATCH=16 timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b16.jsonl 2> gpurun_out/latency_b16.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_n1.json 2> gpurun_out/bench7_n1.err
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vote2 -c 1 \
  -o gpurun_out/vote7_full python benchmarks/profile_step.py 1 > gpurun_out/ncu_vote7.log 2>&1
echo finished > gpurun_out/run7.done
