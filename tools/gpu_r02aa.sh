#!/bin/bash
# share_device mode: the two new tests, then the cohort leg at 1 / 2 / 3 / 4 workers per GPU
set -u
mkdir -p gpurun_out
TAG=r02aa
timeout 600 python -m pytest tests/test_k1_gpu.py tests/test_host_cpp_gpu.py -q -m gpu -k "share or concurrent" > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
nproc
SP_COHORT_DEBUG=1 timeout 300 python tools/cohort_bench.py 6 1 2> gpurun_out/${TAG}_cohort_phases.txt | tail -1; grep "cohort rank" gpurun_out/${TAG}_cohort_phases.txt | tail -4
timeout 900 python tools/cohort_bench.py 48 1 2 3 4 > gpurun_out/${TAG}_cohort_workers.json 2> gpurun_out/${TAG}_cohort_workers.err; cat gpurun_out/${TAG}_cohort_workers.json; tail -3 gpurun_out/${TAG}_cohort_workers.err
