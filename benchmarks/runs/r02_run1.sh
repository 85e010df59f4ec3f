#!/bin/bash
# round 2, GPU run 1: parity of the new voting path + boundary, first bench, vote sweep A/B, ncu of k_vote2
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_backbone.py 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests/test_gpu_backbone.py -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_backbone.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 200 python bench.py --steps 20 --warmup 3 --config 5 --no-cpu-baseline > gpurun_out/bench_n1_cfg5.json 2> gpurun_out/bench_n1_cfg5.err
for impl in 0 2; do
  for hpl in 4 8; do
    [ "$impl" = 0 ] && [ "$hpl" = 8 ] && continue
    PVNET_VOTE_IMPL=$impl PVNET_VOTE_HPL=$hpl SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 \
      python benchmarks/vote_sweep.py > gpurun_out/sweep_impl${impl}_hpl${hpl}.jsonl 2> gpurun_out/sweep_impl${impl}_hpl${hpl}.err
  done
done
timeout 300 python benchmarks/vote_sweep.py > gpurun_out/vote_sweep_full.jsonl 2> gpurun_out/vote_sweep_full.err
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vote2 -c 2 \
  -o gpurun_out/vote2_full python benchmarks/profile_step.py 1 > gpurun_out/ncu_vote.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches.csv python benchmarks/profile_step.py 2 > gpurun_out/ncu_list.log 2>&1
echo finished > gpurun_out/run1.done
