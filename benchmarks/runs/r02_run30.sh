#!/bin/bash
# round 2, GPU run 30: k_vote3 with the bands in shared memory; HPL 8 at 3 CTAs per SM (80 registers, 6 warps per sub-partition)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
PVNET_VOTE_CTAS=3 timeout 900 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py -m gpu -q -rf --tb=short 2>&1 | tail -8 > gpurun_out/pytest_vote3_c3.log
rm -f gpurun_out/vote_sustained.jsonl gpurun_out/vote_sustained.err
for ctas in 2 3; do
  for field in planted random; do
    echo "# ctas $ctas" >> gpurun_out/vote_sustained.jsonl
    PVNET_VOTE_CTAS=$ctas SUST_FIELD=$field SUST_SKIP_BURST=1 timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
  done
done
for ctas in 2 3; do
  PVNET_VOTE_CTAS=$ctas SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 python benchmarks/vote_sweep.py > gpurun_out/sweep30_c$ctas.jsonl 2> gpurun_out/sweep30_c$ctas.err
done
PVNET_VOTE_CTAS=3 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench30_c3.json 2> gpurun_out/bench30_c3.err
echo finished > gpurun_out/run30.done
