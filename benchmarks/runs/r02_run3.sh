#!/bin/bash
# round 2, GPU run 3: warp-autonomous k_vote2, VP kernels, PnP; bench
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
for hpl in 4 8; do
  PVNET_VOTE_HPL=$hpl SWEEP_POINTS="10000:512,50000:2048,150000:2048" timeout 200 \
    python benchmarks/vote_sweep.py > gpurun_out/sweep3_hpl${hpl}.jsonl 2> gpurun_out/sweep3_hpl${hpl}.err
done
PVNET_VOTE_HPL=8 PVNET_VOTE_CTAS=1 SWEEP_POINTS="50000:2048,150000:2048" timeout 200 python benchmarks/vote_sweep.py > gpurun_out/sweep3_hpl8_c1.jsonl 2>&1
PVNET_VOTE_HPL=4 PVNET_VOTE_CTAS=2 SWEEP_POINTS="50000:2048,150000:2048" timeout 200 python benchmarks/vote_sweep.py > gpurun_out/sweep3_hpl4_c2.jsonl 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench3_n1.json 2> gpurun_out/bench3_n1.err
PVNET_VOTE_HPL=4 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench3_n1_hpl4.json 2> gpurun_out/bench3_n1_hpl4.err
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vote2 -c 1 \
  -o gpurun_out/vote3_full python benchmarks/profile_step.py 1 > gpurun_out/ncu_vote3.log 2>&1
echo finished > gpurun_out/run3.done
