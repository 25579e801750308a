#!/bin/bash
# bench.py at N GPUs the way the driver launches it; result + stderr under gpurun_out/
set -u
N=${1:-8}; TAG=${2:-r02p}
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read())
print({k:d[k] for k in ("value","n_gpus","ms_per_step","samples_per_s","gpu_launches")})
print("e2e", d["e2e"]["value"]); print("panel", d["panel"]["gcups"], d["panel"]["ms_per_step"]); print("k1", d["roofline"]["k1_ms"], d["roofline"]["frac"])
PY
