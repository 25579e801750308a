"""GPU suite, row N3: K8 (sp_graph_align through the C++ host: VariantGraph::from_reference_variants, graph_edit_distance, the
walk back, graph_alleles) against the oracle (oracle/graph_oracle.py): same graph, same distance, same traversed nodes, same
allele vector -- and the hand-off to K6 (assign_haplotypes_from_alleles) end to end at CYP2D6 size."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))

import graph_oracle as go  # noqa: E402
import starphase_oracle as so  # noqa: E402
from test_graph_cpu import random_variants, rnd, spell  # noqa: E402

pytestmark = pytest.mark.gpu


def host_typing(bb, region_start, variants, seqs, band):
    from pb_starphase_b200 import _starphase_host as host

    gpu = host.GpuAligner(0)
    res, nodes, n2a, sink = host.graph_typing(gpu, bb.decode(), region_start, [(p, r.decode(), a.decode()) for p, r, a in variants],
                                              [s.decode() for s in seqs], band)
    return res, nodes, n2a, sink


def test_graph_and_alignment_vs_oracle_small():
    rng = np.random.default_rng(3)
    for case in range(40):
        bb = rnd(rng, int(rng.integers(5, 60)))
        variants = random_variants(rng, bb, 700, int(rng.integers(0, 6)))
        g = go.build_graph(bb, 700, variants)
        seqs = []
        for _ in range(3):
            s = bytearray(spell(rng, g)[1])
            for _ in range(int(rng.integers(0, 3))):
                if len(s) > 1:
                    q = int(rng.integers(0, len(s)))
                    s[q:q + 1] = [b"", b"N", s[q:q + 1] + b"C"][int(rng.integers(0, 3))]
            seqs.append(bytes(s))
        seqs.append(b"")
        for band in (4, 40):
            res, nodes, n2a, sink = host_typing(bb, 700, variants, seqs, band)
            assert [(n[0], list(n[1]), n[2]) for n in nodes] == [(g.seqs[k], g.preds[k], g.coord[k]) for k in range(len(g.seqs))], case
            assert {k: [tuple(x) for x in v] for k, v in n2a.items()} == g.node_to_alleles and sink == g.sink
            for s, (found, score, trav, alleles) in zip(seqs, res):
                want_score, want_nodes = go.align(g, s, band=band)
                assert found == (want_score < go.INF), (case, band)
                if found:
                    assert score == want_score and list(trav) == want_nodes, (case, band, s)
                    assert list(alleles) == so.alleles_from_traversal(len(variants), want_nodes, g.node_to_alleles)


def test_cyp2d6_sized_typing_and_k6_handoff():
    """A 6.2 kb backbone with 380 variant sites, consensuses carrying a known set of alleles plus HiFi-like errors: distance,
    traversed nodes and allele vectors equal the oracle's; the vectors then go through K6 (assign_haplotypes_from_alleles) and
    come back as the haplotype they were built from."""
    from pb_starphase_b200 import _starphase_host as host
    from pb_starphase_b200 import synth

    rng = np.random.default_rng(4)
    bb = rnd(rng, 6200)
    pos = sorted(int(x) for x in rng.choice(np.arange(20, 6150, 14), size=380, replace=False))
    variants = []
    for k, p in enumerate(pos):
        if k % 5 == 4:
            variants.append((42_000_000 + p, bb[p:p + 3], bb[p:p + 1]))
        elif k % 5 == 3:
            variants.append((42_000_000 + p, bb[p:p + 1], bb[p:p + 1] + b"GA"))
        else:
            variants.append((42_000_000 + p, bb[p:p + 1], bytes([b"ACGT"[(b"ACGT".index(bb[p]) + 1 + k % 3) % 4]])))
    g = go.build_graph(bb, 42_000_000, variants)
    haps = {f"{h + 1}.001": [int(x) for x in (rng.random(len(variants)) < 0.03 * (h + 1))] for h in range(6)}
    seqs, truth = [], []
    for name, vec in haps.items():
        # spell the path that takes ALT where the haplotype says 1
        path, cur = [0], 0
        while cur != g.sink:
            succ = [k for k, pr in enumerate(g.preds) if cur in pr]
            pick = succ[0]
            for k in succ:
                lab = g.node_to_alleles.get(k)
                if lab and all(vec[v] == a for v, a in lab):
                    pick = k
            path.append(pick)
            cur = pick
        clean = b"".join(g.seqs[k] for k in path)
        noisy, _ = synth.hifi_reads(rng, [clean], 1, err=0.001, flank=0, lo=0, hi=1 << 20)
        seqs.append(noisy[0])
        truth.append(name)
    gpu = host.GpuAligner(0)
    res, nodes, n2a, sink = host.graph_typing(gpu, bb.decode(), 42_000_000, [(p, r.decode(), a.decode()) for p, r, a in variants],
                                              [s.decode() for s in seqs], 96)
    vectors = []
    for s, (found, score, trav, alleles) in zip(seqs, res):
        want_score, want_nodes = go.align(g, s, band=96)
        assert found and score == want_score and list(trav) == want_nodes
        assert list(alleles) == so.alleles_from_traversal(len(variants), want_nodes, g.node_to_alleles)
        vectors.append(list(alleles))
    # K6 hand-off: arg-max of (vi_match, all_match) over the haplotype definitions gives back the haplotype each consensus carries
    meta = [(f"rs{k}", k % 7 == 0) for k in range(len(variants))]
    out = host.assign_haplotypes_from_alleles(gpu, vectors, haps, meta, True)
    assert [o[0] for o in out] == truth
    for o, vec in zip(out, vectors):
        want = so.assign_haplotype_from_alleles(vec, haps, [m[0] for m in meta], [m[1] for m in meta], True)
        assert (o[0], tuple(o[2])) == (want[0], want[2])


def test_find_full_type_in_sequences_vs_oracle(oracle):
    """find_full_type_in_sequence end to end (src/cyp2d6/haplotyper.rs:326-361, :371-601): template search, consensus -> backbone
    mapping (K4 + K9), variant graph of the aligned stretch, K8, allele vector, K6 -- the C++ host against the same flow on
    oracle numbers; consensuses built from known haplotypes come back as those haplotypes."""
    import json

    import flow_oracle as fo
    from pb_starphase_b200 import _starphase_host as host
    from pb_starphase_b200 import synth

    c = synth.cyp2d6_diploid_sample(2001)
    templates = [(t, s, q) for (t, s), q in zip(c["template_labels"], c["templates"])]
    d6 = next(q for t, s, q in templates if t == "CYP2D6" and s is None)
    d7 = next(q for t, s, q in templates if t == "CYP2D7" and s is None)
    rng = np.random.default_rng(9)
    start = 42_100_000
    backbone = rnd(rng, 300) + d6 + rnd(rng, 300)
    pos = sorted(int(x) for x in rng.choice(np.arange(60, len(d6) - 60, 16), size=120, replace=False))
    variants = []
    for k, p in enumerate(pos):
        if k % 6 == 5:
            variants.append((start + 300 + p, d6[p:p + 3], d6[p:p + 1]))
        elif k % 6 == 4:
            variants.append((start + 300 + p, d6[p:p + 1], d6[p:p + 1] + b"TC"))
        else:
            variants.append((start + 300 + p, d6[p:p + 1], bytes([b"ACGT"[(b"ACGT".index(d6[p]) + 1 + k % 3) % 4]])))
    haps = {f"{h + 2}.001": [int(x) for x in (rng.random(len(variants)) < 0.03 * (h + 1))] for h in range(3)}
    haps["1.001"] = [0] * len(variants)
    seqs, truth = [], []
    for name, vec in haps.items():
        s = d6
        for (p, r, a), on in sorted(zip(variants, vec), reverse=True):
            if on:
                q = p - start - 300
                assert s[q:q + len(r)] == r
                s = s[:q] + a + s[q + len(r):]
        noisy, _ = synth.hifi_reads(rng, [s], 1, err=0.0005, flank=0, lo=0, hi=1 << 20)
        seqs.append(noisy[0])
        truth.append(("CYP2D6", name))
    seqs += [d7, rnd(rng, 3000), seqs[1][400:5200]]  # a D7 consensus, junk, an incomplete D6 consensus
    meta = [(f"rs{k}", k % 5 == 0) for k in range(len(variants))]
    mapped = [("CYP2D6", None), ("Hybrid", "CYP2D6::CYP2D7::exon9")]
    db = dict(backbone=backbone, backbone_start=start, variants=variants, metadata=meta, haplotype_lookup=haps, mapped_hybrids=mapped)
    want = fo.find_full_type_in_sequences(oracle, templates, seqs, 0.5, True, db, 96)
    gpu = host.GpuAligner(0)
    got = host.find_full_type_in_sequences(gpu, [(t, s, q.decode()) for t, s, q in templates], [s.decode() for s in seqs], 0.5, True,
                                           backbone.decode(), start, [(p, r.decode(), a.decode()) for p, r, a in variants], meta, haps, mapped, 96)
    assert len(got) == len(want) == len(seqs)
    for g, w in zip(got, want):
        if w is None:
            assert g is None
            continue
        assert (g[0], g[1]) == w[0]
        assert (json.loads(g[2]) if g[2] is not None else None) == w[1]
    assert [w[0] for w in want[:len(truth)]] == truth
    assert want[len(truth)][0] == ("CYP2D7", None) and want[len(truth) + 1] is None
