"""The arithmetic identity K1 / K3 / K4's inner step rests on (pb_starphase_b200/csrc/sp_kernels.cuh `column_step`):
with Hyyro's `hin < 0` bit fed to the Myers add as its carry-in, the carry-IN vector of that addition is the horizontal-minus
vector shifted up one row, and the adder's carry between words is the bit the shift would have moved.  Exhaustive over every
reachable state of a small word, against the textbook block step (Myers 1999 / Hyyro 2003)."""
import itertools


def textbook(W, eq, pv, mv, x, y):
    M = (1 << W) - 1
    eqp = eq | y                       # hin < 0: the row above already paid -> Eq bit 0 forced
    t = eqp & pv
    s = (t + pv) & M
    d0 = ((s ^ pv) | eqp | mv) & M
    ph = (mv | ~(d0 | pv)) & M
    mh = pv & d0
    phs, mhs = ((ph << 1) | x) & M, ((mh << 1) | y) & M
    return (mhs | ~(d0 | phs)) & M, phs & d0, ph >> (W - 1), mh >> (W - 1), d0


def kernel_form(W, eq, pv, mv, x, y):
    M = (1 << W) - 1
    t = eq & pv
    full = t + pv + y                  # the carry-in takes the hin < 0 bit
    c = ((full & M) ^ t ^ pv) & M      # carry-in vector = Mh << 1 | y
    d0 = (c | eq | mv) & M
    ph = (mv | ~(d0 | pv)) & M
    phs = ((ph << 1) | x) & M
    return (c | ~(d0 | phs)) & M, phs & d0, ph >> (W - 1), full >> W, d0


def states(W):
    for cells in itertools.product((0, 1, 2), repeat=W):  # vertical delta per row: 0, +1, -1 (never both)
        pv = sum(1 << i for i, v in enumerate(cells) if v == 1)
        mv = sum(1 << i for i, v in enumerate(cells) if v == 2)
        yield pv, mv


def test_single_word_exhaustive():
    W, n = 6, 0
    for pv, mv in states(W):
        for eq in range(1 << W):
            for x, y in ((0, 0), (1, 0), (0, 1)):
                assert kernel_form(W, eq, pv, mv, x, y) == textbook(W, eq, pv, mv, x, y)
                n += 1
    assert n == 3 ** W * 2 ** W * 3


def test_two_words_linked_by_the_adder_carry_only():
    """Two W-bit words of one lane: the textbook step on the 2W-bit vector equals the kernel form word by word, where the only link
    for the minus deltas is the adder's carry (the plus deltas still need their funnel shift)."""
    W = 4
    M = (1 << W) - 1
    for pv, mv in states(2 * W):
        for eq in (0x00, 0xFF, 0x5A, 0xA5, 0x3C, 0x81, 0x7E, 0x10):
            for x, y in ((0, 0), (1, 0), (0, 1)):
                want = textbook(2 * W, eq, pv, mv, x, y)
                lo = kernel_form(W, eq & M, pv & M, mv & M, x, y)
                # word 1: incoming plus bit = top bit of word 0's Ph (the funnel shift), incoming minus bit = word 0's adder carry
                hi = kernel_form(W, eq >> W, pv >> W, mv >> W, lo[2], lo[3])
                got = (lo[0] | hi[0] << W, lo[1] | hi[1] << W, hi[2], hi[3], lo[4] | hi[4] << W)
                assert got == want
