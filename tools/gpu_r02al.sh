#!/bin/bash
# A/B of the delta collection on the FMA pipe (IMAD.HI, -DSP_COLLECT_HI) against the default SHF collection: K1 alone at the three
# lane widths of the bench, then the K1 / K3 / K4 parity tests on the variant
set -u
mkdir -p gpurun_out
export QB_US=13,12,9 QB_ALLELES=3000 QB_READS=512
for v in default hi default hi; do
  if [ $v = default ]; then unset SP_GPU_LIB; else export SP_GPU_LIB=$PWD/variants/libstarphase_gpu_$v.so; fi
  echo "== $v"; timeout 120 python tools/quick_bench.py 2>&1 | grep '"U"' | cut -c1-120
done > gpurun_out/r02al_collect_ab.txt 2>&1
cat gpurun_out/r02al_collect_ab.txt
export SP_GPU_LIB=$PWD/variants/libstarphase_gpu_hi.so
timeout 200 python -m pytest tests/test_k1_gpu.py tests/test_k3_gpu.py tests/test_k4_gpu.py -x -q -m gpu > gpurun_out/r02al_pytest_hi.log 2>&1; tail -3 gpurun_out/r02al_pytest_hi.log
