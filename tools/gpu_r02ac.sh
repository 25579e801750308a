#!/bin/bash
# 2-GPU sanity of the worker-pool cohort leg under torchrun (short main leg, no panel), the way the driver launches bench.py
set -u
mkdir -p gpurun_out
TAG=${1:-r02ac}
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 ${2:---steps 2 --warmup 3 --panel-reads 0 --cohort-samples 30} > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
tail -3 gpurun_out/${TAG}_bench_2gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_2gpu.json").read())
print({k:d[k] for k in ("value","n_gpus","ms_per_step","samples_per_s","gpu_launches")})
print("e2e", d["e2e"]["value"]); print("cohort", d["cohort"]); print("clocks", d["clocks"])
PY
