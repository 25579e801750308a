#!/usr/bin/env python3
"""Scratch measurement on the GPU box: K1 throughput of one lane width against the text-tile capacity, i.e. against the
number of resident CTAs per SM (smaller tiles -> less shared memory -> 3 CTAs instead of 2)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pb_starphase_b200 as sp
from pb_starphase_b200 import synth

ctx = sp.Context(0)
alleles, reads, _ = synth.hla_gene(1, "HLA-B", n_alleles=int(os.environ.get("QB_ALLELES", "3000")), n_reads=int(os.environ.get("QB_READS", "512")))
cells = sum(len(a) for a in alleles) * sum(len(r) for r in reads)
for U in os.environ.get("QB_US", "9,13").split(","):
    os.environ["SP_FORCE_U"] = U
    P = ctx.patterns(alleles)
    for tc in os.environ.get("QB_TCS", "4096,3000,2000,1400,1000").split(","):
        os.environ["SP_FORCE_TC"] = tc
        T = ctx.targets(reads)
        for rep in range(3):
            d = ctx.score_device(T, P, elem_bits=16)
            ms = ctx.last_kernel_ms(0)
            d.close()
        print(json.dumps(dict(U=U, tc=tc, ms=ms, tcups=cells / ms / 1e9)), flush=True)
        T.close()
    P.close()
