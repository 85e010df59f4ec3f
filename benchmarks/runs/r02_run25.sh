#!/bin/bash
# round 2, GPU run 25: k_vote3 with queued exact decisions -- tests, sustained layer, bench
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py tests/test_gpu_reference_layer.py tests/test_gpu_variants.py -m gpu -q -rf --tb=short 2>&1 | tail -15 > gpurun_out/pytest_vote3_queue.log
rm -f gpurun_out/vote_sustained.jsonl gpurun_out/vote_sustained.err
for field in planted random; do
  SUST_FIELD=$field SUST_SKIP_BURST=1 timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
done
SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 python benchmarks/vote_sweep.py > gpurun_out/sweep25.jsonl 2> gpurun_out/sweep25.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench25_n1.json 2> gpurun_out/bench25_n1.err
echo finished > gpurun_out/run25.done
