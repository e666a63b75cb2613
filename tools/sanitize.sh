#!/bin/bash
# compute-sanitizer over the kernel tests at tiny shapes (memcheck + racecheck), ONE tool per gpurun call
# (B200_PROFILING.md: four tools in one call have left GPUs needing a reset).  Usage on the GPU box:
#     tools/sanitize.sh memcheck|racecheck|synccheck|initcheck [pytest -k expression]
# NOTE (round 2): the pool this repo is developed on answers "compute-sanitizer is closed on this pool" - the script is
# kept for pools where it is open.  Pipeline protocol mistakes of the tcgen05 kernels are caught without it by the
# bounded mbarrier waits (dq_la_tc_last_error).
set -e
tool=${1:-memcheck}
expr=${2:-"linear_attention or resnet_block or scheduler or multiplex"}
cd "$(dirname "$0")/.."
export PYTHONPATH=diffusion-deconvolution-dia-msms-data_b200:oracle
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$expr" 2>&1 | tail -60 | tee gpurun_out/sanitize_$tool.log
