#!/bin/bash
# ncu launch list of the bench command at HEAD (main leg), plus the tests added after the r02ae record
set -u
mkdir -p gpurun_out
TAG=r02ag
timeout 300 python -m pytest tests/test_k2_gpu.py -q -m gpu > gpurun_out/${TAG}_pytest_k2.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_k2.log
B="--panel-reads 0 --cohort-samples 0 --cpu-seconds 2"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 $B > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1; head -20 gpurun_out/${TAG}_launches_summary.txt
