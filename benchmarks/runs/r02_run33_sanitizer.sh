#!/bin/bash
# round 2, GPU run 33 (final vote kernel): compute-sanitizer on the new kernels (small test subset: the tools slow kernels 50-100x)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
SEL='not config3 and not config4 and not config5 and not adversarial and not reference_layer'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_pnp.py tests/test_gpu_variants.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_memcheck_vote_pnp.log 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_memcheck_vote_pnp.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "different_thresholds or device_rng_matches or workspace or huge" > gpurun_out/sanitizer_racecheck_vote.log 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_racecheck_vote.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_jpeg.py -m gpu -q -x -k "not dataparallel_two" > gpurun_out/sanitizer_memcheck_boundary.log 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_memcheck_boundary.log
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "different_thresholds or device_rng_matches" > gpurun_out/sanitizer_initcheck_vote.log 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_initcheck_vote.log
echo finished > gpurun_out/run33.done
