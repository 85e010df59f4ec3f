#!/bin/bash
# round 2, GPU run 29 (8 GPUs): the driver's N=8 launch of bench.py, once, to see the line it will get
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo finished > gpurun_out/run29.done
