#!/bin/bash
# compute-sanitizer over the kernels written or rewritten in round 2 (K4, K5 weighted, K7, K8, comm world 1), small cases
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02u_sanitizer.txt
: > $OUT
SEL="tests/test_k4_gpu.py tests/test_k5_gpu.py tests/test_consensus_gpu.py::test_k7_extend_vs_oracle tests/test_graph_gpu.py::test_graph_and_alignment_vs_oracle_small tests/test_comm_gpu.py::test_world1_comm_equals_plain_calls tests/test_k3_gpu.py"
SP_SKIP_LARGE=1 timeout 1500 compute-sanitizer --tool memcheck --target-processes all python -m pytest $SEL -x -q -m gpu -k "not class_ii and not long_text" 2>&1 | tail -6 >> $OUT
echo "memcheck rc=$?" >> $OUT
SP_SKIP_LARGE=1 timeout 1500 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_k4_gpu.py::test_align_small_known tests/test_k4_gpu.py::test_align_vs_oracle_edge_lengths tests/test_consensus_gpu.py::test_k7_extend_vs_oracle tests/test_graph_gpu.py::test_graph_and_alignment_vs_oracle_small -x -q -m gpu 2>&1 | tail -6 >> $OUT
echo "racecheck rc=$?" >> $OUT
cat $OUT
