#!/usr/bin/env python3
"""Scratch timing on the GPU box of the C++ host's CYP2D6 template search (Cyp2d6Extractor::find_base_type_in_sequences)
on one full-size diploid sample; SP_TIMING=1 prints the phases of each sp_align_pairs call."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pb_starphase_b200 import _starphase_host as host
from pb_starphase_b200 import synth

gpu = host.GpuAligner(0)
c = synth.cyp2d6_diploid_sample(2001)
templates = [(t, s, q.decode()) for (t, s), q in zip(c["template_labels"], c["templates"])]
reads = [r.decode() for r in c["reads"]]
for rep in range(3):
    t0 = time.perf_counter()
    hits = host.find_base_type_in_sequences(gpu, templates, reads, False, 0.5)
    print(f"find_base_type_in_sequences: {1e3 * (time.perf_counter() - t0):.1f} ms, {sum(map(len, hits))} hits", flush=True)
