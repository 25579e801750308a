#!/usr/bin/env python3
"""Scratch measurement on the GPU box: what the affine refinement (K9) costs on top of the unit-cost placement (K4), at the
score_read shape (every allele of a gene, DNA and cDNA, against one consensus) through the C++ host's align_pairs, and for
the host flows that use it (template search of one CYP2D6 sample).  Writes gpurun_out/k9_bench.json."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pb_starphase_b200 import _starphase_host as host
from pb_starphase_b200 import synth

out = {}
gpu = host.GpuAligner(0)
n_alleles = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
alleles, reads, src, cdna = synth.hla_gene(3, "HLA-A", n_alleles=n_alleles, n_reads=64, with_cdna=True)
pats = [a.decode() for a in list(alleles) + list(cdna)]
targets = [reads[0].decode(), cdna[int(src[0])].decode()]
pairs = [(0, p) for p in range(n_alleles)] + [(1, n_alleles + p) for p in range(n_alleles)]


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return best, r


for refine in (False, True):
    host.set_stand_ins(affine_refine=refine)
    dt, res = timed(lambda: gpu.align_pairs(targets, pats, pairs, 5))
    key = "score_read_shape_k4_k9" if refine else "score_read_shape_k4_only"
    out[key] = dict(pairs=len(pairs), ms=1e3 * dt, refined=sum(1 for r in res if r["refined"]),
                    nm_sum=sum(r["nm"] for r in res), score_sum=sum(r["score"] for r in res))
    print(key, json.dumps(out[key]), flush=True)

c = synth.cyp2d6_diploid_sample(2001)
templates = [(t, s, q.decode()) for (t, s), q in zip(c["template_labels"], c["templates"])]
creads = [r.decode() for r in c["reads"]]
for refine in (False, True):
    host.set_stand_ins(affine_refine=refine)
    dt, hits = timed(lambda: host.find_base_type_in_sequences(gpu, templates, creads, False, 0.5))
    key = "template_search_k4_k9" if refine else "template_search_k4_only"
    out[key] = dict(ms=1e3 * dt, hits=sum(map(len, hits)))
    print(key, json.dumps(out[key]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/k9_bench.json").write_text(json.dumps(out, indent=1))
