"""GPU suite, row N1: K7 (sp_consensus_extend, through the C ABI) against the oracle's banded recurrence, bit for bit, and the
C++ host search (ConsensusDWFA / DualConsensusDWFA above K7) against the oracle's search: same consensus sequences, same per-read
scores, same read split.  Parity with waffle_con itself is unpinned; the properties of tests/test_consensus_cpu.py
are re-checked here at HLA size."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))

import consensus_oracle as co  # noqa: E402
from pb_starphase_b200 import synth  # noqa: E402
from test_consensus_cpu import het_pair, rnd  # noqa: E402

pytestmark = pytest.mark.gpu


def test_k7_extend_vs_oracle(ctx):
    import pb_starphase_b200 as sp

    rng = np.random.default_rng(5)
    src = rnd(rng, 180)
    reads, _ = synth.hifi_reads(rng, [src], 9, err=0.02, flank=0, lo=0, hi=1 << 20)
    reads += [src[60:], src[75:160], b"", b"ACGTNACGT", src[:50], b"*" * 30 + src[30:], src[:90] + b"**" + src[92:]]  # '*' = wildcard
    offsets = [-1] * 9 + [60, 80, -1, 10, 0, -1, -1]
    for band, window in ((8, 0), (32, 0), (16, 30), (100, 64)):
        cfg = co.Config(band=band, offset_window=window)
        cons = sp.Consensus(ctx, reads, offsets, offset_window=window, band=band, max_tracks=6)
        tracks = {0: co.Track(len(reads))}
        cons.reset(0)
        ed, votes, full = cons.extend([0], [0], [0])
        tracks[0], e0, v0, f0 = co.extend(reads, offsets, cfg, tracks[0], 0)  # the report also folds column 0 into best_full
        assert ed[0].tolist() == e0 and votes[0].tolist() == v0 and full[0].tolist() == f0
        # walk the true sequence with a detour: two children per step (the right symbol and a wrong one), parent = the right child
        cur = 0
        for pos, base in enumerate(src[:150]):
            wrong = b"ACGT"[(b"ACGT".index(base) + 1 + pos % 3) % 4]
            free = [t for t in range(6) if t != cur][:2]
            ed, votes, full = cons.extend([cur, cur], [base, wrong], free)
            for q, symbol in enumerate((base, wrong)):
                t, e, v, f = co.extend(reads, offsets, cfg, tracks[cur], symbol)
                assert ed[q].tolist() == e, (band, window, pos, q)
                assert votes[q].tolist() == v and full[q].tolist() == f, (band, window, pos, q)
                tracks[free[q]] = t
            cur = free[0]
        # in place (src == dst), and a report of the result
        ed, votes, full = cons.extend([cur], [src[150]], [cur])
        t, e, v, f = co.extend(reads, offsets, cfg, tracks[cur], src[150])
        assert ed[0].tolist() == e and votes[0].tolist() == v and full[0].tolist() == f
        ed, votes, full = cons.extend([cur], [0], [cur])
        assert ed[0].tolist() == e and votes[0].tolist() == v and full[0].tolist() == f
        with pytest.raises(sp.SpError):
            cons.extend([0, 1], [65, 65], [2, 2])  # two tasks writing one track
        cons.close()


def host_consensus(reads, offsets=None, **cfg):
    from pb_starphase_b200 import _starphase_host as host

    gpu = host.GpuAligner(0)
    offs = list(offsets) if offsets is not None else []
    return gpu, host, [r.decode() for r in reads], offs, cfg


@pytest.mark.parametrize("on_device", [True, False])
def test_host_search_equals_oracle_search(on_device, monkeypatch):
    """on_device: stretches with one way forward run on the device (sp_consensus_run); off: every symbol is stepped from the host
    (SP_CONSENSUS_NO_RUN).  Same answers either way, and the same as the oracle's search."""
    if not on_device:
        monkeypatch.setenv("SP_CONSENSUS_NO_RUN", "1")
    rng = np.random.default_rng(6)
    for case in range(4):
        a, b = het_pair(rng, 260 + 20 * case, (30, 131, 222))
        na, nb = int(rng.integers(4, 9)), int(rng.integers(0, 7))
        ra, _ = synth.hifi_reads(rng, [a], na, err=0.006, flank=0, lo=0, hi=1 << 20)
        rb, _ = synth.hifi_reads(rng, [b], nb, err=0.006, flank=0, lo=0, hi=1 << 20) if nb else ([], None)
        reads = ra + rb
        gpu, host, rs, offs, _ = host_consensus(reads)
        got, calls1 = host.consensus(gpu, rs, offs, {})
        want = co.consensus(reads)
        assert [(s, list(sc)) for s, sc in got] == [(s, sc) for s, sc in want], case
        gotd, calls = host.dual_consensus(gpu, rs, offs, {})
        wantd = co.dual_consensus(reads)
        assert len(gotd) == len(wantd)
        for g, w in zip(gotd, wantd):
            assert g["consensus1"] == w["consensus1"] and g["consensus2"] == w["consensus2"], case
            assert list(g["is_consensus1"]) == w["is_consensus1"] and list(g["scores1"]) == w["scores1"] and list(g["scores2"]) == w["scores2"], case
        print(f"case {case}: on_device={on_device} device calls: single {calls1}, dual {calls}")
        if not on_device:
            assert calls > 200  # one device call per expanded node


def test_host_search_offsets_and_windows():
    rng = np.random.default_rng(7)
    src = rnd(rng, 420)
    reads = [src[:300], src[:330], src[:310], src[100:], src[125:], src[90:], src]
    offsets = [None, None, None, 112, 120, 90, None]
    for window in (40, 400):  # 400: the reference's setting (src/hla/caller.rs:1113), the widest band class
        cfg = dict(allow_early_termination=True, offset_window=window)
        gpu, host, rs, offs, _ = host_consensus(reads, offsets)
        got, _ = host.consensus(gpu, rs, offs, cfg)
        want = co.consensus(reads, offsets, co.Config(**cfg))
        assert [(s, list(sc)) for s, sc in got] == want and got[0][0] == src
        gotd, _ = host.dual_consensus(gpu, rs, offs, cfg)
        wantd = co.dual_consensus(reads, offsets, co.Config(**cfg))
        assert [(g["consensus1"], g["consensus2"], list(g["scores1"])) for g in gotd] == [(w["consensus1"], w["consensus2"], w["scores1"]) for w in wantd]


def test_hla_sized_dual_consensus_properties():
    """30 HiFi-like reads of two 3.3 kb alleles differing at five sites: both alleles come back exactly, every read on its side."""
    rng = np.random.default_rng(8)
    a, b = het_pair(rng, 3300, (200, 900, 1700, 2500, 3100))
    ra, _ = synth.hifi_reads(rng, [a], 16, err=0.002, flank=0, lo=0, hi=1 << 20)
    rb, _ = synth.hifi_reads(rng, [b], 14, err=0.002, flank=0, lo=0, hi=1 << 20)
    reads = ra + rb
    gpu, host, rs, offs, _ = host_consensus(reads)
    import time

    t0 = time.perf_counter()
    got, calls = host.dual_consensus(gpu, rs, offs, {})
    print(f"dual consensus of 30 reads x 3.3 kb: {1e3 * (time.perf_counter() - t0):.1f} ms, {calls} device calls")
    d = got[0]
    assert {d["consensus1"], d["consensus2"]} == {a, b}
    first_is_a = d["consensus1"] == a
    assert list(d["is_consensus1"]) == [first_is_a] * 16 + [not first_is_a] * 14
    single, _ = host.consensus(gpu, [r.decode() for r in ra], [], {})
    assert single[0][0] == a


@pytest.mark.parametrize("hpc_only", [False, True])
def test_hla_consensus_step_vs_oracle(hpc_only):
    """run_dual_consensus_with_offsets + the per-group re-consensus (src/hla/caller.rs:1151-1219, :706-760) through the C++ host
    against the same flow on the oracle's search.  hpc_only: the alleles differ only in a homopolymer length, so the
    homopolymer-compressed pass sees one allele and the full-length DNA pass has to split the reads."""
    import flow_oracle as fo
    import starphase_oracle as so
    from pb_starphase_b200 import _starphase_host as host

    rng = np.random.default_rng(12 + hpc_only)
    a = rnd(rng, 800)  # the differences lie behind the offset window of the partial reads (400, src/hla/caller.rs:1113)
    if hpc_only:
        q = 600
        b = a[:q] + a[q:q + 1] * 2 + a[q:]  # one homopolymer two bases longer
    else:
        b = bytearray(a)
        for q in (520, 640, 730):
            b[q] = b"ACGT"[(b"ACGT".index(bytes([a[q]])) + 1) % 4]
        b = bytes(b)
    records = []
    for k in range(10):
        src = a if k % 2 == 0 else b
        start = 0 if k < 6 else int(rng.integers(20, 120))
        end = len(src) if k % 3 else len(src) - int(rng.integers(0, 40))
        seq = src[start:end]
        records.append((f"read{k:02d}", seq, so.hpc(seq), start, so.hpc_pos(src, start)))
    want_d, want_pass, want_groups = fo.hla_consensus_step(records)
    gpu = host.GpuAligner(0)
    got_d, got_pass, got_groups = host.hla_consensus_step(gpu, [(q, s.decode(), h.decode(), o, ho) for q, s, h, o, ho in records], host.DiplotypeSettings())
    assert got_d["consensus1"] == want_d["consensus1"] and got_d["consensus2"] == want_d["consensus2"]
    assert list(got_d["is_consensus1"]) == want_d["is_consensus1"] and list(got_d["scores1"]) == want_d["scores1"] and list(got_d["scores2"]) == want_d["scores2"]
    assert got_pass == want_pass and tuple(got_groups) == want_groups
    assert want_pass and set(want_groups) == {a, b}
    assert want_d["is_consensus1"] in ([k % 2 == 0 for k in range(10)], [k % 2 == 1 for k in range(10)])  # the reads split by source allele


def priority_case(rng):
    """Four sources: A, B (two SNVs: differs at both levels), C (a homopolymer of A two bases longer: same HPC as A), D (unrelated,
    seeded apart); five reads each, chain = (homopolymer-compressed, raw)."""
    import starphase_oracle as so

    a = rnd(rng, 300)
    b = bytearray(a)
    for q in (80, 190):
        b[q] = b"ACGT"[(b"ACGT".index(bytes([a[q]])) + 2) % 4]
    b = bytes(b)
    c = a[:150] + a[150:151] * 2 + a[150:]
    d = rnd(rng, 260)
    sources = [a, b, c, d]
    chains, offsets, seeds, truth = [], [], [], []
    for k in range(20):
        src = sources[k % 4]
        chains.append([so.hpc(src), src])
        offsets.append([None, None])
        seeds.append(3 if k % 4 == 3 else None)
        truth.append(k % 4)
    return sources, chains, offsets, seeds, truth


def test_priority_consensus_vs_oracle():
    """PriorityConsensusDWFA (src/cyp2d6/caller.rs:145-280) through the C++ host against the oracle's restatement, and the property
    that pins both: the reads come back grouped by source, every group with its (HPC, raw) consensus pair."""
    import starphase_oracle as so
    from pb_starphase_b200 import _starphase_host as host

    sources, chains, offsets, seeds, truth = priority_case(np.random.default_rng(21))
    want_cons, want_idx = co.priority_consensus(chains, offsets, seeds)
    gpu = host.GpuAligner(0)
    got_cons, got_idx = host.priority_consensus(gpu, [[x.decode() for x in ch] for ch in chains], offsets, seeds, {})
    assert list(got_idx) == want_idx
    assert [[(s, list(sc)) for s, sc in levels] for levels in got_cons] == [[(s, list(sc)) for s, sc in levels] for levels in want_cons]
    assert len(want_cons) == 4
    for k, g in enumerate(want_idx):
        assert want_cons[g][1][0] == sources[truth[k]] and want_cons[g][0][0] == so.hpc(sources[truth[k]])
