#!/bin/bash
# share_device experiments: what the priority side stream and the one-CTA-per-item launch each contribute (3 workers, 36 samples)
set -u
mkdir -p gpurun_out
TAG=r02ab
{
echo "# A: three workers, share_device off (persistent K1 grids of three contexts side by side)"
SP_COHORT_NOSHARE=1 timeout 300 python tools/cohort_bench.py 36 3 2>/dev/null
echo "# B: three workers, share_device on, long K1 launches on the main stream too (SP_SHARE_NOBULK=1)"
SP_SHARE_NOBULK=1 timeout 300 python tools/cohort_bench.py 36 3 2>/dev/null
echo "# C: three workers, share_device on (default)"
timeout 300 python tools/cohort_bench.py 36 3 2>/dev/null
echo "# D: six workers, share_device on"
timeout 300 python tools/cohort_bench.py 36 6 2>/dev/null
} > gpurun_out/${TAG}_share_experiments.txt
cut -c1-120 gpurun_out/${TAG}_share_experiments.txt
SP_COHORT_DEBUG=1 timeout 300 python tools/cohort_bench.py 12 3 2> gpurun_out/${TAG}_phases_3workers.txt >/dev/null; grep "cohort rank" gpurun_out/${TAG}_phases_3workers.txt | tail -12
