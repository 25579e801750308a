"""GPU parity test for K5 (sp_row_topk: the k best patterns of every target, the stand-in for minimap2's best_n hit
list at src/hla/realigner.rs:116-146), through the C ABI, against a stable numpy sort of the same matrix."""
import numpy as np
import pytest

from test_k3_gpu import noisy, rnd

pytestmark = pytest.mark.gpu


def expect(D, k):
    order = np.argsort(D, axis=1, kind="stable")[:, :k]
    idx = np.full((D.shape[0], k), -1, dtype=np.int32)
    dist = np.full((D.shape[0], k), -1, dtype=np.int32)
    idx[:, :order.shape[1]] = order
    dist[:, :order.shape[1]] = np.take_along_axis(D, order, axis=1)
    return idx, dist


@pytest.mark.parametrize("bits", [16, 32])
def test_row_topk_vs_numpy(ctx, bits):
    rng = np.random.default_rng(21)
    pats = [rnd(rng, int(m)) for m in rng.integers(40, 400, 150)]
    pats += [pats[3], pats[3], pats[77]]                      # exact duplicates: ties resolved by the lower index
    texts = [rnd(rng, 10) + noisy(rng, pats[i % len(pats)], int(rng.integers(0, 9))) + rnd(rng, 12) for i in range(300)]
    T, P = ctx.targets(texts), ctx.patterns(pats)
    d = ctx.score_device(T, P, elem_bits=bits)
    D = (d.to_host_u16() if bits == 16 else d.to_host()).astype(np.int64)
    for k in (1, 5, 8, 16):
        idx, dist = ctx.row_topk(d, k)
        eidx, edist = expect(D, k)
        assert (idx == eidx).all() and (dist == edist).all(), k
    # sp_row_topk_biased with bias = max |P| - |P|: "most pattern bases explained first"; dist stays the plain distance
    lens = np.array([len(p) for p in pats], dtype=np.int64)
    bias = (lens.max() - lens).astype(np.int32)
    for k in (1, 5, 16):
        idx, dist = ctx.row_topk(d, k, bias=bias)
        order = np.argsort(D + bias[None, :], axis=1, kind="stable")[:, :k]
        assert (idx == order).all() and (dist == np.take_along_axis(D, order, axis=1)).all(), k
    # sp_row_topk_weighted: key = 5 * distance + bias, the realigner's stand-in for minimap2's score ranking
    for k in (1, 5, 16):
        idx, dist = ctx.row_topk(d, k, bias=bias, weight=5)
        order = np.argsort(5 * D + bias[None, :], axis=1, kind="stable")[:, :k]
        assert (idx == order).all() and (dist == np.take_along_axis(D, order, axis=1)).all(), k
    import pb_starphase_b200 as sp

    with pytest.raises(sp.SpError):
        ctx.row_topk(d, 5, bias=-np.ones(len(pats), dtype=np.int32))
    with pytest.raises(sp.SpError):
        ctx.row_topk(d, 5, bias=bias, weight=65)
    d.close(); T.close(); P.close()


def test_row_topk_fewer_patterns_than_k(ctx):
    d = ctx.score_device(ctx.targets([b"ACGTACGT", b"TTTT", b""]), ctx.patterns([b"ACGT", b"TT"]), elem_bits=16)
    idx, dist = ctx.row_topk(d, 5)
    assert idx.tolist() == [[0, 1, -1, -1, -1], [1, 0, -1, -1, -1], [1, 0, -1, -1, -1]]
    assert dist.tolist() == [[0, 1, -1, -1, -1], [0, 3, -1, -1, -1], [2, 4, -1, -1, -1]]
    import pb_starphase_b200 as sp

    with pytest.raises(sp.SpError):
        ctx.row_topk(d, 17)
