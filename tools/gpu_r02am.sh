#!/bin/bash
# the GPU suite at the final commit of round 2
set -u
mkdir -p gpurun_out
timeout 280 python -m pytest tests -q -m gpu > gpurun_out/r02am_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02am_pytest_gpu.log
