#!/bin/bash
# round 2, GPU run 26: the state to be judged -- full gpu tests, smoke, bench (both arms), ncu tables
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench26_n1.json 2> gpurun_out/bench26_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench26_reference.json 2> gpurun_out/bench26_reference.err
timeout 300 python bench.py --config 5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench26_cfg5.json 2> gpurun_out/bench26_cfg5.err
timeout 300 python bench.py --config 4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench26_cfg4.json 2> gpurun_out/bench26_cfg4.err
timeout 300 python benchmarks/vote_sweep.py > gpurun_out/vote_sweep_final.jsonl 2> gpurun_out/vote_sweep_final.err
timeout 300 python benchmarks/latency_b1.py > gpurun_out/latency_b1.jsonl 2> gpurun_out/latency_b1.err
timeout 300 python benchmarks/step_breakdown.py > gpurun_out/step_breakdown.txt 2> gpurun_out/step_breakdown.err
BREAKDOWN_COV=1 timeout 300 python benchmarks/step_breakdown.py >> gpurun_out/step_breakdown.txt 2>> gpurun_out/step_breakdown.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_final.csv python benchmarks/profile_step.py 2 > gpurun_out/ncu_list.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vote3 -c 1 \
  -o gpurun_out/vote_final python benchmarks/profile_step.py 1 > gpurun_out/ncu_vote_final.log 2>&1
timeout 600 python benchmarks/gpu_baselines.py > gpurun_out/gpu_baselines.jsonl 2> gpurun_out/gpu_baselines.err
SUST_FIELD=planted timeout 200 python benchmarks/vote_sustained.py > gpurun_out/vote_sustained_final.jsonl 2> gpurun_out/vote_sustained_final.err
SUST_FIELD=random timeout 200 python benchmarks/vote_sustained.py >> gpurun_out/vote_sustained_final.jsonl 2>> gpurun_out/vote_sustained_final.err
echo finished > gpurun_out/run26.done
