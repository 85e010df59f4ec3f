#!/bin/bash
# round 2, GPU run 36: full gpu test suite on the committed state (with the graph-mode pipeline test)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo finished > gpurun_out/run36.done
