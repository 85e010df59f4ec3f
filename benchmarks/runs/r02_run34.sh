#!/bin/bash
# round 2, GPU run 34: fused-head pixel-major records staged through shared memory (coalesced stores)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_boundary.py tests/test_gpu_pipeline.py tests/test_gpu_conv.py -m gpu -q -rf --tb=short 2>&1 | tail -12 > gpurun_out/pytest_headstage.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench34_n1.json 2> gpurun_out/bench34_n1.err
timeout 300 python benchmarks/step_breakdown.py > gpurun_out/step_breakdown34.txt 2> gpurun_out/step_breakdown34.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches34.csv python benchmarks/profile_step.py 2 > gpurun_out/ncu_list34.log 2>&1
echo finished > gpurun_out/run34.done
