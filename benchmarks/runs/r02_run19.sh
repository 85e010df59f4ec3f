#!/bin/bash
# round 2, GPU run 19: k_vote3 form 3 (packed count adds), HPL 4 x 3 CTAs, at 16 items per warp
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/vote_sustained.jsonl gpurun_out/vote_sustained.err
PVNET_VOTE_FORM=3 timeout 600 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py -m gpu -q -rf --tb=short 2>&1 | tail -8 > gpurun_out/pytest_vote3_form3.log
for cfg in "2 8 2 4" "3 8 2 4" "3 8 2 8" "2 4 3 4" "3 4 3 4" "3 4 4 4"; do
  set -- $cfg
  for field in planted random; do
    echo "# form $1 hpl $2 ctas $3 group $4" >> gpurun_out/vote_sustained.jsonl
    PVNET_VOTE_FORM=$1 PVNET_VOTE_HPL=$2 PVNET_VOTE_CTAS=$3 PVNET_VOTE_GROUP=$4 SUST_FIELD=$field SUST_SKIP_BURST=1 timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
  done
done
echo finished > gpurun_out/run19.done
