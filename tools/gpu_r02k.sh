#!/bin/bash
# round 2, call k: smoke, full-size default bench (N=1) + reference arm, ncu launch list of the same command
set -u
mkdir -p gpurun_out
TAG=r02k
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
SP_TIMING=1 timeout 300 python tools/host_bench.py 2>&1 | tail -3
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_reference.json
