#!/bin/bash
# ncu --set full of K1 after the 7 + 4 inner step (reduced scale: a 2 s launch replayed ~45 times does not fit a call)
set -u
mkdir -p gpurun_out
TAG=r02aj
B="--panel-reads 0 --cohort-samples 0 --cpu-seconds 1"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k1_infix -s 2 -c 1 -o gpurun_out/${TAG}_k1_full -f \
    python bench.py --steps 1 --warmup 1 --scale 0.05 $B > gpurun_out/${TAG}_ncu_k1.log 2>&1
python tools/ncu_summary.py full gpurun_out/${TAG}_k1_full.ncu-rep > gpurun_out/${TAG}_k1_ncu_full.txt 2>&1; rm -f gpurun_out/${TAG}_k1_full.ncu-rep
cat gpurun_out/${TAG}_k1_ncu_full.txt | head -50
