#!/bin/bash
# round 2, GPU run 32 (4 GPUs): the driver's N=8 launch of bench.py, once, to see the line it will get
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus4.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo finished > gpurun_out/run32.done
