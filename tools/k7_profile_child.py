#!/usr/bin/env python3
"""One HLA-sized single consensus (16 reads x 3.3 kb): the workload of the ncu capture of k7_run."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

from pb_starphase_b200 import _starphase_host as host
from pb_starphase_b200 import synth
from test_consensus_gpu import het_pair

rng = np.random.default_rng(8)
a, b = het_pair(rng, 3300, (200, 900, 1700, 2500, 3100))
ra, _ = synth.hifi_reads(rng, [a], 16, err=0.002, flank=0, lo=0, hi=1 << 20)
gpu = host.GpuAligner(0)
out, calls = host.consensus(gpu, [r.decode() for r in ra], [], {})
print(len(out[0][0]), calls)
