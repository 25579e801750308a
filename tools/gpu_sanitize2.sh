#!/bin/bash
# compute-sanitizer over what came after r02u: K9 (affine), k7_run (the on-device consensus loop, priority chain), derive_texts,
# K1 in share_device mode (one CTA per item, two contexts from two threads), three aligners on one GPU
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02ad_sanitizer.txt
: > $OUT
SEL="tests/test_affine_gpu.py tests/test_consensus_gpu.py::test_host_search_offsets_and_windows tests/test_consensus_gpu.py::test_priority_consensus_vs_oracle tests/test_k1_gpu.py::test_targets_derive_splice_and_revcomp tests/test_k1_gpu.py::test_share_device_mode_is_bit_identical tests/test_host_cpp_gpu.py::test_concurrent_aligners_share_one_gpu"
SP_SKIP_LARGE=1 timeout 1500 compute-sanitizer --tool memcheck --target-processes all python -m pytest $SEL -x -q -m gpu 2>&1 | tail -6 >> $OUT
echo "memcheck rc=$?" >> $OUT
SP_SKIP_LARGE=1 timeout 1200 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_affine_gpu.py::test_affine_small_full_band tests/test_affine_gpu.py::test_affine_windows_and_nothing_to_align tests/test_consensus_gpu.py::test_host_search_offsets_and_windows tests/test_k1_gpu.py::test_targets_derive_splice_and_revcomp -x -q -m gpu 2>&1 | tail -6 >> $OUT
echo "racecheck rc=$?" >> $OUT
cat $OUT
