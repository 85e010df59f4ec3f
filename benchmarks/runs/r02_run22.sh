#!/bin/bash
# round 2, GPU run 22: k_vote3 with the count taken from the sign bit (LEA.HI) -- tests, sustained layer, sweep
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py tests/test_gpu_reference_layer.py tests/test_gpu_variants.py -m gpu -q -rf --tb=short 2>&1 | tail -15 > gpurun_out/pytest_vote3_sign.log
rm -f gpurun_out/vote_sustained.jsonl gpurun_out/vote_sustained.err
for grp in 4 8; do
  for field in planted random; do
    echo "# group $grp" >> gpurun_out/vote_sustained.jsonl
    PVNET_VOTE_GROUP=$grp SUST_FIELD=$field SUST_SKIP_BURST=1 timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
  done
done
SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 python benchmarks/vote_sweep.py > gpurun_out/sweep22.jsonl 2> gpurun_out/sweep22.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench22_n1.json 2> gpurun_out/bench22_n1.err
echo finished > gpurun_out/run22.done
