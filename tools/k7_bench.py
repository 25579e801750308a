#!/usr/bin/env python3
"""Scratch timing on the GPU box: the consensus search (K7) at HLA size -- 30 HiFi-like reads of two 3.3 kb alleles -- single and
dual, with the on-device run (sp_consensus_run) and stepped from the host (SP_CONSENSUS_NO_RUN=1), for 1..8 CTAs per cluster
(SP_K7_RUN_CTAS)."""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np

    from pb_starphase_b200 import _starphase_host as host
    from pb_starphase_b200 import synth
    from test_consensus_gpu import het_pair

    rng = np.random.default_rng(8)
    a, b = het_pair(rng, 3300, (200, 900, 1700, 2500, 3100))
    ra, _ = synth.hifi_reads(rng, [a], 16, err=0.002, flank=0, lo=0, hi=1 << 20)
    rb, _ = synth.hifi_reads(rng, [b], 14, err=0.002, flank=0, lo=0, hi=1 << 20)
    gpu = host.GpuAligner(0)
    # a second of K1 first: a lone small kernel on an idle GPU runs at idle clocks, which is not how the search runs inside a sample
    alleles, wreads, _ = synth.hla_gene(3, "HLA-A", n_alleles=3000, n_reads=256)
    t0 = time.perf_counter()
    gpu.score_batch([r.decode() for r in wreads], [x.decode() for x in alleles])
    print(f"  warm-up K1: {1e3 * (time.perf_counter() - t0):.0f} ms", flush=True)
    for name, reads, fn in (("single 16 reads", ra, host.consensus), ("dual 30 reads", ra + rb, host.dual_consensus)):
        rs = [r.decode() for r in reads]
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            out, calls = fn(gpu, rs, [], {})
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        print(f"  {name}: {1e3 * best:.1f} ms, {calls} device calls", flush=True)
    sys.exit(0)

for env in ({"SP_CONSENSUS_NO_RUN": "1"}, {"SP_K7_RUN_CTAS": "1"}, {"SP_K7_RUN_CTAS": "2"}, {"SP_K7_RUN_CTAS": "4"}, {"SP_K7_RUN_CTAS": "8"}, {}):
    print(env or "default", flush=True)
    subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), check=False)
