#!/bin/bash
# round 2, GPU run 35b: the state to be judged after the upsampling work -- full gpu tests, smoke, bench (both arms,
# configs 2/4/5), step breakdown, ncu launch list
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench35_n1.json 2> gpurun_out/bench35_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench35_reference.json 2> gpurun_out/bench35_reference.err
timeout 300 python bench.py --config 4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench35_cfg4.json 2> gpurun_out/bench35_cfg4.err
timeout 300 python bench.py --config 5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench35_cfg5.json 2> gpurun_out/bench35_cfg5.err
timeout 300 python benchmarks/step_breakdown.py > gpurun_out/step_breakdown.txt 2> gpurun_out/step_breakdown.err
BREAKDOWN_COV=1 timeout 300 python benchmarks/step_breakdown.py >> gpurun_out/step_breakdown.txt 2>> gpurun_out/step_breakdown.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_final.csv python benchmarks/profile_step.py 2 > gpurun_out/ncu_list.log 2>&1
echo finished > gpurun_out/run35.done
