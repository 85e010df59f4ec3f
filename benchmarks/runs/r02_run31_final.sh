#!/bin/bash
# round 2, last GPU run: the committed state once more -- gpu tests, smoke, both bench arms
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench31_reference.json 2> gpurun_out/bench31_reference.err
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench31_n1.json 2> gpurun_out/bench31_n1.err
echo finished > gpurun_out/run31.done
