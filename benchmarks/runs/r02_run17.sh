#!/bin/bash
# round 2, GPU run 17: k_vote3 arithmetic forms (0 edges+FMNMX3, 1 num-|perp| + two FMNMX, 2 num-|perp| + FMNMX3)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/vote_sustained.jsonl gpurun_out/vote_sustained.err
for form in 1 2; do
  PVNET_VOTE_FORM=$form timeout 600 python -m pytest tests/test_gpu_vote.py tests/test_gpu_pipeline.py -m gpu -q -rf --tb=short 2>&1 | tail -8 > gpurun_out/pytest_vote3_form$form.log
done
for form in 0 1 2; do
  for grp in 4 8; do
    for field in planted random; do
      echo "# form $form group $grp" >> gpurun_out/vote_sustained.jsonl
      PVNET_VOTE_FORM=$form PVNET_VOTE_GROUP=$grp SUST_FIELD=$field SUST_SKIP_BURST=1 timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
    done
  done
done
echo finished > gpurun_out/run17.done
