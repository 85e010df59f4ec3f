#!/bin/bash
# round 2, GPU run 18: k_vote3 form 2, items per warp sweep; bench
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/vote_sustained.jsonl gpurun_out/vote_sustained.err
for items in 6 10 16 24; do
  for field in planted random; do
    echo "# items $items" >> gpurun_out/vote_sustained.jsonl
    PVNET_VOTE_ITEMS=$items SUST_FIELD=$field SUST_SKIP_BURST=1 timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained.jsonl 2>> gpurun_out/vote_sustained.err
  done
done
for items in 6 12 24; do
  PVNET_VOTE_ITEMS=$items SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 python benchmarks/vote_sweep.py > gpurun_out/sweep18_items$items.jsonl 2> gpurun_out/sweep18_items$items.err
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench18_n1.json 2> gpurun_out/bench18_n1.err
PVNET_VOTE_ITEMS=16 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench18_n1_items16.json 2> gpurun_out/bench18_n1_items16.err
echo finished > gpurun_out/run18.done
