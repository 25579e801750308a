#!/usr/bin/env python3
"""Cohort leg of bench.py alone (BASELINE configs[4]) at several worker counts per GPU: samples/s with 1 .. W host threads,
one sp_ctx each (sp_ctx_share_device for W > 1).  usage: cohort_bench.py <samples> <workers> [<workers> ...]"""
import hashlib
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

samples = int(sys.argv[1]) if len(sys.argv) > 1 else 40
counts = [int(x) for x in sys.argv[2:]] or [1, 2, 3]
t0 = time.perf_counter()
w = bench.build_workload(1, 1.0)
print(f"workload built in {time.perf_counter() - t0:.1f} s", file=sys.stderr, flush=True)
for n in counts:
    dt, ns, nbytes, info = bench.run_cohort(w, samples, 0, 1, 0, n)
    print(json.dumps(dict(workers=n, samples=ns, samples_per_s=ns / dt, ms_per_sample=1e3 * dt / ns, json_bytes=nbytes,
                          calls=info["last_sample_calls"], launches=info["kernel_launches"])), flush=True)
