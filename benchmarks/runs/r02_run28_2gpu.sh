#!/bin/bash
# round 2, GPU run 28 (2 GPUs, final kernels): DataParallel test over two devices, N=2 bench under torchrun (both arms)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_jpeg.py -m gpu -q -rf --tb=short 2>&1 | tail -20 > gpurun_out/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_n2_reference.json 2> gpurun_out/bench_n2_reference.err
echo finished > gpurun_out/run28.done
