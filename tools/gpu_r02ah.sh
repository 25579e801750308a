#!/bin/bash
# the Mhs = carry-in form of the Myers step (7 ALU-pipe + 4 FMA-pipe instructions per word): GPU suite, then the main leg
set -u
mkdir -p gpurun_out
TAG=r02ah
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --panel-reads 0 --cohort-samples 0 --cpu-seconds 1 > gpurun_out/${TAG}_bench_main.json 2> gpurun_out/${TAG}_bench_main.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_main.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "k1_ms", d["roofline"]["k1_ms"], "k1_tcups", d["roofline"]["k1_tcups"], "frac", d["roofline"]["frac"], "best", d["best_pairs"])
PY
