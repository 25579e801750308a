#!/bin/bash
# round 2, call a: HEAD sanity (GPU tests), K4 / K3 baselines with phase timing, first ncu captures of k4_align and k3_span_starts
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02a_pytest_gpu.log
SP_TIMING=1 timeout 300 python tools/k4_bench.py > gpurun_out/r02a_k4_bench.log 2>&1; tail -30 gpurun_out/r02a_k4_bench.log
timeout 300 python tools/span_bench.py > gpurun_out/r02a_span_bench.log 2>&1; tail -3 gpurun_out/r02a_span_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_align -s 1 -c 1 -o gpurun_out/r02a_k4_full -f python tools/k4_bench.py 1000 > gpurun_out/r02a_ncu_k4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_span_starts -s 1 -c 1 -o gpurun_out/r02a_k3_full -f python tools/span_bench.py > gpurun_out/r02a_ncu_k3.log 2>&1
ls -la gpurun_out | tail -8
