#!/usr/bin/env python3
"""Scratch measurements on the GPU box (not the bench contract): INT pipe peaks and K1 GCUPS per lane width."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import pb_starphase_b200 as sp
from pb_starphase_b200 import synth

out = {}
ctx = sp.Context(0)
for kind, name in enumerate(["lop3", "imad", "lop3+imad", "iadd", "imadhi", "lop3+imadhi"]):
    out[f"int_peak_{name}_Tops"] = ctx.int_peak(kind) / 1e12
print(json.dumps(out), flush=True)

n_alleles = int(os.environ.get("QB_ALLELES", "2000"))
n_reads = int(os.environ.get("QB_READS", "256"))
alleles, reads, _ = synth.hla_gene(1, "HLA-B", n_alleles=n_alleles, n_reads=n_reads)
T = ctx.targets(reads)
cells = sum(len(a) for a in alleles) * sum(len(r) for r in reads)
for U in os.environ.get("QB_US", "4,6,8,10,12,16,auto").split(","):
    if U == "auto":
        os.environ.pop("SP_FORCE_U", None)
    else:
        os.environ["SP_FORCE_U"] = U
    P = ctx.patterns(alleles)
    for rep in range(3):
        d = ctx.score_device(T, P, elem_bits=16)
        ms = ctx.last_kernel_ms(0)
        d.close()
    rec = dict(U=U, ms=ms, tcups=cells / ms / 1e9, padded_frac=P.total_len / P.padded_rows)
    out[f"k1_U{U}"] = rec
    print(json.dumps(rec), flush=True)
    P.close()
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/quick_bench.json").write_text(json.dumps(out, indent=1))
