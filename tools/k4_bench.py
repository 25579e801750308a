#!/usr/bin/env python3
"""Scratch measurement on the GPU box: K4 (sp_align_pairs) at the score_read shape -- every allele of a gene (DNA and
cDNA) against one consensus -- and at the realign shape (many reads x 5 candidates)."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import pb_starphase_b200 as sp
from pb_starphase_b200 import synth

out = {}
ctx = sp.Context(0)
n_alleles = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
alleles, reads, src, cdna = synth.hla_gene(3, "HLA-A", n_alleles=n_alleles, n_reads=64, with_cdna=True)
pats = list(alleles) + list(cdna)
targets = [reads[0], cdna[int(src[0])]]
pairs = [(0, p) for p in range(n_alleles)] + [(1, n_alleles + p) for p in range(n_alleles)]
for rep in range(2):
    t0 = time.perf_counter()
    res = ctx.align_pairs(targets, pats, pairs)
    dt = time.perf_counter() - t0
cells = sum(len(pats[p]) * len(targets[t]) for t, p in pairs)
out["score_read_shape"] = dict(pairs=len(pairs), seconds=dt, kernel_ms=ctx.last_kernel_ms(4), pairs_per_s=len(pairs) / dt, gcups_forward=2 * cells / dt / 1e9)
print(json.dumps(out["score_read_shape"]), flush=True)
# realign shape: 64 reads x 5 candidate alleles
D = ctx.score_batch(reads, alleles)
cand = np.argsort(D, axis=1, kind="stable")[:, :5]
pairs = [(r, int(a)) for r in range(len(reads)) for a in cand[r]]
for rep in range(2):
    t0 = time.perf_counter()
    res = ctx.align_pairs(reads, alleles, pairs)
    dt = time.perf_counter() - t0
out["realign_shape"] = dict(pairs=len(pairs), seconds=dt, kernel_ms=ctx.last_kernel_ms(4), pairs_per_s=len(pairs) / dt)
print(json.dumps(out["realign_shape"]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/k4_bench.json").write_text(json.dumps(out, indent=1))
