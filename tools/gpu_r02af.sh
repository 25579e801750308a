#!/bin/bash
# main leg only, twice: does `value` meet `e2e` now that the clock sampler polls once a second without power.draw?
set -u
mkdir -p gpurun_out
for k in 1 2; do
timeout 600 python bench.py --panel-reads 0 --cohort-samples 0 --cpu-seconds 1 > gpurun_out/r02af_bench_main_$k.json 2> gpurun_out/r02af_bench_main_$k.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02af_bench_main_$k.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "k1_ms", d["roofline"]["k1_ms"], "clocks", d["clocks"])
PY
done
