#!/bin/bash
# round 2 profiles: ncu launch list of the bench command, K1 DRAM traffic at full scale, full captures of K1 / K2 / K4 / K3 spans / K7
set -u
mkdir -p gpurun_out
TAG=r02s
B="--panel-reads 0 --cohort-samples 0 --cpu-seconds 2"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 $B > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k1_infix -c 12 --csv \
    --log-file gpurun_out/${TAG}_k1_traffic.csv python bench.py --steps 1 --warmup 1 $B > gpurun_out/${TAG}_ncu_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_infix -s 2 -c 1 -o gpurun_out/${TAG}_k1_full -f \
    python bench.py --steps 1 --warmup 1 --scale 0.05 $B > gpurun_out/${TAG}_ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_pair -s 2 -c 1 -o gpurun_out/${TAG}_k2_full -f \
    python bench.py --steps 1 --warmup 1 --scale 0.25 $B > gpurun_out/${TAG}_ncu_k2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_align -s 2 -c 1 -o gpurun_out/${TAG}_k4_full -f \
    python tools/k4_bench.py 2000 > gpurun_out/${TAG}_ncu_k4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_span_starts -s 1 -c 1 -o gpurun_out/${TAG}_k3_full -f \
    python tools/span_bench.py > gpurun_out/${TAG}_ncu_k3.log 2>&1
timeout 300 python tools/k4_bench.py > gpurun_out/${TAG}_k4_bench.log 2>&1; grep pairs gpurun_out/${TAG}_k4_bench.log
timeout 300 python tools/span_bench.py 2>&1 | tail -1
ls -la gpurun_out | grep ${TAG}
