"""CPU suite: the variant-graph oracle (oracle/graph_oracle.py, row N3) holds the properties that pin it -- hiphase's WFAGraph
cannot be run here (parity unpinned): the DP equals brute force over every path on small graphs; a sequence spelled from the
backbone with chosen alleles has distance 0 and traverses exactly those allele nodes; indistinguishable alleles come out
ambiguous (2) through alleles_from_traversal (src/cyp2d6/haplotyper.rs:452-468)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))

import graph_oracle as go  # noqa: E402
import starphase_oracle as so  # noqa: E402


def rnd(rng, n):
    return bytes(rng.choice(list(b"ACGT"), n).tolist())


def random_variants(rng, bb: bytes, region_start: int, n_var: int):
    vs = []
    n = len(bb)
    for _ in range(n_var):
        p = int(rng.integers(0, n))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            vs.append((region_start + p, bb[p:p + 1], bytes([b"ACGT"[(b"ACGT".index(bb[p]) + 1 + int(rng.integers(0, 3))) % 4]])))
        elif kind == 1 and p + 3 <= n:
            vs.append((region_start + p, bb[p:p + 3], bb[p:p + 1]))      # deletion, VCF style (anchor base kept)
        else:
            vs.append((region_start + p, bb[p:p + 1], bb[p:p + 1] + rnd(rng, 2)))  # insertion
    return vs


def spell(rng, g: go.Graph):
    path = [0]
    while path[-1] != g.sink:
        path.append(int(rng.choice([k for k, pr in enumerate(g.preds) if path[-1] in pr])))
    return path, b"".join(g.seqs[k] for k in path)


def test_dp_equals_brute_force_on_small_graphs():
    rng = np.random.default_rng(0)
    for case in range(150):
        bb = rnd(rng, int(rng.integers(5, 36)))
        g = go.build_graph(bb, 1000, random_variants(rng, bb, 1000, int(rng.integers(0, 5))))
        assert all(g.seqs[k] for k in g.node_to_alleles)
        _, seq = spell(rng, g)
        seq = bytearray(seq)
        for _ in range(int(rng.integers(0, 3))):
            if len(seq) > 1:
                q = int(rng.integers(0, len(seq)))
                seq[q:q + 1] = [b"", bytes([b"ACGT"[(b"ACGT".index(seq[q]) + 1) % 4]]), seq[q:q + 1] + b"G"][int(rng.integers(0, 3))]
        assert go.align(g, bytes(seq)) == go.brute_force(g, bytes(seq)), case


def test_spelled_path_is_recovered_and_typed():
    rng = np.random.default_rng(1)
    bb = rnd(rng, 900)
    pos = sorted(int(x) for x in rng.choice(np.arange(10, 880, 12), size=30, replace=False))
    variants = []
    for k, p in enumerate(pos):
        if k % 3 == 0:
            variants.append((5000 + p, bb[p:p + 1], bytes([b"ACGT"[(b"ACGT".index(bb[p]) + 2) % 4]])))
        elif k % 3 == 1:
            variants.append((5000 + p, bb[p:p + 4], bb[p:p + 1]))
        else:
            variants.append((5000 + p, bb[p:p + 1], bb[p:p + 1] + b"TTGA"))
    variants.append((9000, b"A", b"C"))  # outside the region: stays unset
    g = go.build_graph(bb, 5000, variants)
    path, seq = spell(rng, g)
    score, nodes = go.align(g, seq, band=64)
    assert score == 0
    labelled = sorted(k for k in path if k in g.node_to_alleles)
    assert [k for k in nodes if k in g.node_to_alleles] == labelled
    alleles = so.alleles_from_traversal(len(variants), nodes, g.node_to_alleles)
    want = [3] * len(variants)
    for k in labelled:
        for v, a in g.node_to_alleles[k]:
            want[v] = a
    assert alleles == want and alleles[-1] == 3 and set(alleles[:-1]) <= {0, 1}


def test_indistinguishable_alleles_are_ambiguous():
    # the sequence stops being informative at the site: deleting the variant base from the read makes REF and ALT equally good
    bb = b"ACGTACGTTTGACCAGTACCGGTTAACGT"
    variants = [(100 + 12, b"C", b"G")]
    g = go.build_graph(bb, 100, variants)
    seq = bb[:12] + bb[13:]
    score, nodes = go.align(g, seq)
    assert score == 1
    assert so.alleles_from_traversal(1, nodes, g.node_to_alleles) == [2]
    assert so.alleles_from_traversal(1, go.align(g, bb)[1], g.node_to_alleles) == [0]
    assert so.alleles_from_traversal(1, go.align(g, bb[:12] + b"G" + bb[13:])[1], g.node_to_alleles) == [1]
