#!/bin/bash
# round 2, GPU run 41b (the round's last GPU seconds): smoke on the rebuilt library, then the dedicated-warp fused upsampling with
# the head MMA two tiles behind (one CTA per SM: the MMA warp no longer waits for the previous tile's epilogue)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke41.log 2>&1
timeout 60 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short -k "dedicated" 2>&1 | tail -6 > gpurun_out/pytest_dedicated41.log
PVNET_FUSE_UP=2 timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench41_up2.json 2> gpurun_out/bench41_up2.err
echo finished > gpurun_out/run41.done
