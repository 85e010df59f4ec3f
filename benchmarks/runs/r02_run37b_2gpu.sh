#!/bin/bash
# round 2, GPU run 37b (2 GPUs): nn.DataParallel boundary tests on the committed state, 2-rank bench line
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_dropin.py -m gpu -q -rf --tb=short 2>&1 | tail -12 > gpurun_out/pytest_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench37_n2.json 2> gpurun_out/bench37_n2.err
echo finished > gpurun_out/run37.done
