#!/bin/bash
# round 2, final record: GPU test suite, smoke, K9 capture, the default bench line
set -u
mkdir -p gpurun_out
TAG=r02y
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k9_affine_local -c 1 -o gpurun_out/${TAG}_k9_full -f \
    python tools/k9_bench.py 1500 > gpurun_out/${TAG}_ncu_k9.log 2>&1
ncu -i gpurun_out/${TAG}_k9_full.ncu-rep --page details > gpurun_out/${TAG}_k9_ncu_details.txt 2>/dev/null; rm -f gpurun_out/${TAG}_k9_full.ncu-rep
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json
ls -la gpurun_out | grep ${TAG}
