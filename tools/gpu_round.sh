#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the same command and one
# `ncu --set full` capture of K1.  Everything lands under gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_reference.json
# launch list of the same command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 2 > gpurun_out/${TAG}_ncu_bench.log 2>&1
# one full capture of K1 on a scaled allele set (40 replays of a 5 s kernel would eat the budget)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_infix -s 2 -c 1 \
    -o gpurun_out/${TAG}_k1_full -f python bench.py --steps 1 --warmup 1 --scale 0.05 --cpu-seconds 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_pair -s 2 -c 1 \
    -o gpurun_out/${TAG}_k2_full -f python bench.py --steps 1 --warmup 1 --scale 0.25 --cpu-seconds 1 > gpurun_out/${TAG}_ncu_full_k2.log 2>&1
ls -la gpurun_out | tail -20
