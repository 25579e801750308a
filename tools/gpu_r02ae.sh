#!/bin/bash
# round 2, closing record: GPU test suite, racecheck of the consensus loop after the __syncwarp fix, smoke, the default bench line
set -u
mkdir -p gpurun_out
TAG=${1:-r02ae}
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_consensus_gpu.py::test_host_search_offsets_and_windows tests/test_consensus_gpu.py::test_priority_consensus_vs_oracle -x -q -m gpu 2>&1 | tail -8 > gpurun_out/${TAG}_racecheck_k7.txt; cat gpurun_out/${TAG}_racecheck_k7.txt | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-300
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","samples_per_s","gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "panel", d["panel"]["gcups"], "clocks", d["clocks"])
print("cohort", {k: d["cohort"][k] for k in ("samples_per_s","workers_per_gpu","ms_per_sample_per_gpu")}, "cpu", d["cpu_baseline"]["value"])
PY
