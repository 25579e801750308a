"""Test-side loader of the CPU oracle (oracle/sp_oracle.c).  Only tests, smoke() and bench.py's
cpu_baseline / --impl reference legs import this; the product never does."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Sequence

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB = ORACLE_DIR / "libsp_oracle.so"


class PairRec(C.Structure):
    _fields_ = [("score", C.c_uint64), ("score2", C.c_uint64), ("i", C.c_uint32), ("j", C.c_uint32), ("c1", C.c_uint32), ("pad", C.c_uint32)]


def build_oracle(force: bool = False) -> Path:
    newest = max((ORACLE_DIR / f).stat().st_mtime for f in ("sp_oracle.c", "sp_oracle_affine.c", "Makefile"))
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        subprocess.check_call(["make", "-C", str(ORACLE_DIR), "-B" if force else "-s"])
    return LIB


def _pack(seqs: Sequence[bytes]):
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        np.cumsum([len(s) for s in seqs], out=offs[1:])
    joined = b"".join(bytes(s) for s in seqs)
    bases = np.frombuffer(joined, dtype=np.uint8).copy() if joined else np.zeros(1, dtype=np.uint8)
    return bases, offs


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(str(build_oracle()))
        L = self.lib
        L.sp_oracle_infix_dp.restype = C.c_int64
        L.sp_oracle_infix_dp.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
        L.sp_oracle_infix_myers.restype = C.c_int64
        L.sp_oracle_infix_myers.argtypes = L.sp_oracle_infix_dp.argtypes
        L.sp_oracle_score_batch.restype = C.c_int64
        L.sp_oracle_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.sp_oracle_pair_minsum_topk.restype = C.c_int
        L.sp_oracle_pair_minsum_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(PairRec)]
        L.sp_oracle_pair_minsum_full.restype = None
        L.sp_oracle_pair_minsum_full.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        L.sp_oracle_num_threads.restype = C.c_int
        L.sp_oracle_chain_windows.restype = None
        L.sp_oracle_chain_windows.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.sp_oracle_span_batch.restype = None
        L.sp_oracle_span_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]

    def align(self, pattern: bytes, text: bytes) -> dict:
        """Canonical traceback alignment (sp_oracle_align): the same dict Context.align_pairs returns per pair."""
        L = self.lib
        L.sp_oracle_align.restype = C.c_int64
        L.sp_oracle_align.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
        rec = np.zeros(7, dtype=np.int32)
        cap = len(pattern) + len(text) + 2
        cig = np.zeros(cap, dtype=np.uint32)
        n = L.sp_oracle_align(bytes(pattern), len(pattern), bytes(text), len(text), rec.ctypes.data, cig.ctypes.data, cap)
        assert n >= 0
        return {"dist": int(rec[0]), "nm": int(rec[1]), "p_start": int(rec[2]), "p_end": int(rec[3]), "t_start": int(rec[4]),
                "t_end": int(rec[5]), "cigar": [(int(x) >> 4, int(x) & 15) for x in cig[:n]]}

    def num_threads(self) -> int:
        return int(self.lib.sp_oracle_num_threads())

    def infix(self, pattern: bytes, text: bytes, prefix: bool = False, impl: str = "dp"):
        e = C.c_int64(0)
        fn = self.lib.sp_oracle_infix_dp if impl == "dp" else self.lib.sp_oracle_infix_myers
        d = fn(bytes(pattern), len(pattern), bytes(text), len(text), int(prefix), C.byref(e))
        return int(d), int(e.value)

    def score_batch(self, targets, patterns, prefix: bool = False, impl: str = "myers", nthreads: int = 0,
                    want_end_col: bool = False):
        """Returns D[t, p] int32 (and end columns) plus the cell count via .last_cells."""
        tb, to = targets if isinstance(targets, tuple) else _pack(targets)
        pb, po = patterns if isinstance(patterns, tuple) else _pack(patterns)
        nt, npat = len(to) - 1, len(po) - 1
        D = np.zeros((nt, npat), dtype=np.int32)
        E = np.zeros((nt, npat), dtype=np.int32) if want_end_col else None
        self.last_cells = int(self.lib.sp_oracle_score_batch(
            tb.ctypes.data, to.ctypes.data, nt, pb.ctypes.data, po.ctypes.data, npat, int(prefix),
            0 if impl == "dp" else 1, nthreads, D.ctypes.data, E.ctypes.data if E is not None else None))
        return (D, E) if want_end_col else D

    def score_spans(self, targets, patterns, nthreads: int = 0):
        """(D, start, end), each [nt, np] int32: see sp_oracle_span."""
        tb, to = targets if isinstance(targets, tuple) else _pack(targets)
        pb, po = patterns if isinstance(patterns, tuple) else _pack(patterns)
        nt, npat = len(to) - 1, len(po) - 1
        D, S, E = (np.zeros((nt, npat), dtype=np.int32) for _ in range(3))
        self.lib.sp_oracle_span_batch(tb.ctypes.data, to.ctypes.data, nt, pb.ctypes.data, po.ctypes.data, npat, nthreads,
                                      D.ctypes.data, S.ctypes.data, E.ctypes.data)
        return D, S, E

    def chain_windows(self, chains, read_weights, n_haps: int) -> np.ndarray:
        """B[r, c]: see sp_oracle_chain_windows."""
        coff = np.zeros(len(chains) + 1, dtype=np.int32)
        np.cumsum([len(c) for c in chains], out=coff[1:])
        items = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.int32) for c in chains] + [np.zeros(1, np.int32)]), dtype=np.int32)
        soff = np.zeros(len(read_weights) + 1, dtype=np.int32)
        np.cumsum([len(w) for w in read_weights], out=soff[1:])
        W = np.ascontiguousarray(np.concatenate([np.asarray(w, dtype=np.uint32).reshape(-1, n_haps) for w in read_weights] + [np.zeros((1, n_haps), np.uint32)]), dtype=np.uint32)
        B = np.zeros((len(read_weights), len(chains)), dtype=np.int32)
        self.lib.sp_oracle_chain_windows(len(chains), coff.ctypes.data, items.ctypes.data, len(read_weights), soff.ctypes.data,
                                         W.ctypes.data, n_haps, B.ctypes.data)
        return B

    def pair_minsum_topk(self, D: np.ndarray, k: int, nthreads: int = 0, D2=None):
        D = np.ascontiguousarray(D, dtype=np.int32)
        D2 = np.ascontiguousarray(D2, dtype=np.int32) if D2 is not None else None
        recs = (PairRec * k)()
        n = self.lib.sp_oracle_pair_minsum_topk(D.ctypes.data, D2.ctypes.data if D2 is not None else None,
                                                D.shape[0], D.shape[1], k, nthreads, recs)
        if D2 is None:
            return [(int(r.score), int(r.i), int(r.j), int(r.c1)) for r in recs[:n]]
        return [(int(r.score), int(r.score2), int(r.i), int(r.j), int(r.c1)) for r in recs[:n]]

    def pair_minsum_full(self, D: np.ndarray, nthreads: int = 0) -> np.ndarray:
        D = np.ascontiguousarray(D, dtype=np.int32)
        S = np.zeros((D.shape[1], D.shape[1]), dtype=np.uint64)
        self.lib.sp_oracle_pair_minsum_full(D.ctypes.data, D.shape[0], D.shape[1], nthreads, S.ctypes.data)
        return S


# minimap2 cost sets at the reference's call sites: {a, b, q, e, q2, e2}
COSTS_MAP_HIFI = (1, 4, 6, 2, 26, 1)      # standard_hifi_aligner(), src/util/mapping.rs:8-14 (realigner, CYP2D6, consensus -> reference)
COSTS_ALLELE_SCORING = (5, 4, 6, 2, 26, 1)  # mapopt.a = 5 in score_read, src/hla/caller.rs:1370-1381


class AffineOracle:
    """Second-opinion oracle (oracle/sp_oracle_affine.c): best local alignment under minimap2's two-piece affine cost model.
    `.align(pattern, text)` returns the dict shape of Oracle.align plus `score`, so tests/flow_oracle.py runs unchanged on it;
    `prefetch` fills the cache for many pairs with OpenMP over pairs."""

    def __init__(self, costs=COSTS_ALLELE_SCORING):
        self.lib = C.CDLL(str(build_oracle()))
        self.costs = np.asarray(costs, dtype=np.int32)
        L = self.lib
        L.sp_oracle_affine_local.restype = C.c_int64
        L.sp_oracle_affine_local.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.sp_oracle_affine_local_banded.restype = C.c_int64
        L.sp_oracle_affine_local_banded.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                                    C.c_void_p, C.c_int64]
        L.sp_oracle_affine_batch.restype = C.c_int64
        L.sp_oracle_affine_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        self.cache = {}

    @staticmethod
    def _rec(rec, cig):
        return {"dist": int(rec[7]), "nm": int(rec[1]), "p_start": int(rec[2]), "p_end": int(rec[3]), "t_start": int(rec[4]),
                "t_end": int(rec[5]), "cigar": [(int(x) >> 4, int(x) & 15) for x in cig], "score": int(rec[0])}

    def align(self, pattern: bytes, text: bytes, centre: int = 0, band: int = -1) -> dict:
        """band < 0: the unbanded model; else the DP restricted to |(j - i) - centre| <= band, as the product's K9 computes it."""
        key = (bytes(pattern), bytes(text)) if band < 0 else (bytes(pattern), bytes(text), centre, band)
        hit = self.cache.get(key)
        if hit is not None:
            return hit
        rec = np.zeros(8, dtype=np.int32)
        cap = len(pattern) + len(text) + 2
        cig = np.zeros(cap, dtype=np.uint32)
        n = self.lib.sp_oracle_affine_local_banded(key[0], len(key[0]), key[1], len(key[1]), self.costs.ctypes.data, centre, band,
                                                   rec.ctypes.data, cig.ctypes.data, cap)
        assert n >= 0
        out = self.cache[key] = self._rec(rec, cig[:n])
        return out

    def align_batch(self, targets, patterns, pairs, nthreads: int = 0, want_cigar: bool = True):
        """pairs = [(target index, pattern index)]; returns the list of dicts (cigar = None when not wanted)."""
        tb, to = _pack(targets)
        pb, po = _pack(patterns)
        pt = np.ascontiguousarray([t for t, _ in pairs], dtype=np.int32)
        pp = np.ascontiguousarray([p for _, p in pairs], dtype=np.int32)
        n = len(pairs)
        recs = np.zeros((n, 8), dtype=np.int32)
        offs = np.zeros(n, dtype=np.int64)
        cap = int(sum(min(len(patterns[p]) + len(targets[t]) + 2, 4096) for t, p in pairs)) if want_cigar else 0
        cig = np.zeros(max(cap, 1), dtype=np.uint32)
        self.lib.sp_oracle_affine_batch(tb.ctypes.data, to.ctypes.data, pb.ctypes.data, po.ctypes.data, n, pt.ctypes.data, pp.ctypes.data,
                                        self.costs.ctypes.data, nthreads, recs.ctypes.data, offs.ctypes.data,
                                        cig.ctypes.data if want_cigar else None, cap)
        out = []
        for k in range(n):
            if want_cigar:
                assert offs[k] >= 0, "affine CIGAR pool overflow"
                out.append(self._rec(recs[k], cig[offs[k]:offs[k] + recs[k][6]]))
            else:
                d = self._rec(recs[k], [])
                d["cigar"] = None
                out.append(d)
        return out

    def prefetch(self, patterns, text: bytes, nthreads: int = 0):
        """Aligns every pattern against one text in parallel and keeps the results for .align()."""
        todo = [p for p in dict.fromkeys(bytes(p) for p in patterns) if (p, bytes(text)) not in self.cache]
        for p, r in zip(todo, self.align_batch([text], todo, [(0, k) for k in range(len(todo))], nthreads)):
            self.cache[(p, bytes(text))] = r
