#!/usr/bin/env python3
"""Scratch measurement on the GPU box: K3 spans (sp_score_spans_filtered) at the weight_sequence shape of one
CYP2D6 sample (~650 read segments x 24 consensuses), wall time per call."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pb_starphase_b200 as sp
from pb_starphase_b200 import synth

ctx = sp.Context(0)
c = synth.cyp2d6_sample(1000)
out = {}
for rep in range(3):
    t0 = time.perf_counter()
    D, S, E = ctx.score_spans(c["consensuses"], c["segments"], max_dist_permille=350)
    dt = time.perf_counter() - t0
cells = sum(map(len, c["consensuses"])) * sum(map(len, c["segments"]))
out["weight_sequence_shape"] = dict(segments=len(c["segments"]), consensuses=len(c["consensuses"]), seconds=dt,
                                    k1_ms=ctx.last_kernel_ms(0), gcups_forward=cells / dt / 1e9, reported=int((S >= 0).sum()))
print(json.dumps(out), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/span_bench.json").write_text(json.dumps(out, indent=1))
