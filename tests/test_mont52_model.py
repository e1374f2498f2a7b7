"""CPU model of the FP64-pipe Montgomery multiplication of paillier_b200/csrc/mont52.cuh (no GPU needed).

The kernel's arithmetic is exact by construction -- fma_rz(a, b, 2^104) and fma_rz(a, b, 2^104 + 2^52 - hi) return the high and
low 52 bits of a 104-bit product as doubles whose BIT PATTERNS 0x467<<52 | hi and 0x433<<52 | lo are accumulated as 64-bit integers.
What can go wrong is the bookkeeping around it: the constants that cancel the exponent fields (column initial values, the spill
column's start value, the final per-position correction), the one-column shift between lanes, the carry out of column 0 and the
three-pass normalisation.  This model runs exactly that bookkeeping, lane by lane with 64-bit wrap-around, for every built shape
and checks r = a*b*R^-1 mod n with r < 2n, for random and extreme operands."""
import random

import pytest

M64 = (1 << 64) - 1
M52 = (1 << 52) - 1
BH, BL = 0x46700000, 0x43300000
ROWB = (2 * BL + 2 * BH) & 0xFFFFFFFF

SHAPES = [(4, 5, 32), (4, 10, 64), (8, 5, 64), (4, 15, 96), (8, 8, 96), (8, 10, 128), (8, 15, 192), (16, 8, 192)]   # PGPU_FOR_EACH_SHAPE52


def split52(a, b):
    p = a * b
    assert a <= M52 and b <= M52
    return ((0x467 << 52) | (p >> 52)) & M64, ((0x433 << 52) | (p & M52)) & M64


def normalize(C, tpi, L):
    """Mont52::normalize: in-lane ripple, carry-outs to the next lane, second ripple, ballot look-ahead, third ripple"""
    lanes = [C[t * L:(t + 1) * L] for t in range(tpi)]
    outs = []
    for ln in lanes:
        c = 0
        for k in range(L):
            v = (ln[k] + c) & M64
            ln[k], c = v & M52, v >> 52
        outs.append(c)
    assert all(o < (1 << 12) for o in outs)
    c2s = []
    for t, ln in enumerate(lanes):
        c2 = outs[t - 1] if t else 0
        for k in range(L):
            v = ln[k] + c2
            ln[k], c2 = v & M52, v >> 52
        c2s.append(c2)
    g = sum(1 << t for t in range(tpi) if c2s[t])
    p = sum(1 << t for t in range(tpi) if all(x == M52 for x in lanes[t]))
    assert g & p == 0
    ci = ((g | p) + g) ^ p
    for t, ln in enumerate(lanes):
        c3 = (ci >> t) & 1
        for k in range(L):
            v = ln[k] + c3
            ln[k], c3 = v & M52, v >> 52
    return [x for ln in lanes for x in ln]


def mont52_mul(a, b, n, np_, tpi, L):
    """Mont52::mul on limb lists (little-endian, tpi*L limbs of 52 bits); returns the normalised limbs of the result"""
    A = [a[t * L:(t + 1) * L] for t in range(tpi)]
    N = [n[t * L:(t + 1) * L] for t in range(tpi)]
    C = [[((-(k * ROWB + 2 * BL)) & 0xFFFFFFFF) << 32 for k in range(L)] + [0] for _ in range(tpi)]
    KSP = ((-(L * ROWB)) & 0xFFFFFFFF) << 32
    for u in range(tpi):
        for k in range(L):
            bj = b[u * L + k]
            for t in range(tpi):
                h, l = zip(*[split52(A[t][i], bj) for i in range(L)])
                C[t][0] = (C[t][0] + l[0]) & M64
                for i in range(1, L):
                    C[t][i] = (C[t][i] + l[i] + h[i - 1]) & M64
                C[t][L] = (KSP + h[L - 1]) & M64
            q = ((C[0][0] & M52) * np_) & M52
            for t in range(tpi):
                h, l = zip(*[split52(N[t][i], q) for i in range(L)])
                C[t][0] = (C[t][0] + l[0]) & M64
                for i in range(1, L):
                    C[t][i] = (C[t][i] + l[i] + h[i - 1]) & M64
                C[t][L] = (C[t][L] + h[L - 1]) & M64
            assert C[0][0] & M52 == 0                     # column 0 is a multiple of 2^52 and clean of exponent fields
            sends = [C[t][0] for t in range(tpi)]
            for t in range(tpi):
                recv = sends[t + 1] if t + 1 < tpi else 0
                carry = sends[0] >> 52 if t == 0 else 0
                C[t] = C[t][1:L] + [(C[t][L] + recv) & M64, 0]
                C[t][0] = (C[t][0] + carry) & M64
    X = []
    for t in range(tpi):
        for k in range(L):
            X.append((C[t][k] + ((((k + 1) * ROWB - 2 * BH) & 0xFFFFFFFF) << 32)) & M64)
    assert all(x < (1 << 63) for x in X)                  # every exponent field is gone: plain column sums
    return normalize(X, tpi, L)


def limbs(x, count):
    return [(x >> (52 * i)) & M52 for i in range(count)]


def value(ls):
    return sum(v << (52 * i) for i, v in enumerate(ls))


@pytest.mark.parametrize("tpi,L,S32", SHAPES)
def test_mont52_bookkeeping(tpi, L, S32):
    rnd = random.Random(52 * tpi + L)
    s52 = tpi * L
    bits = 32 * S32
    assert 52 * s52 >= bits + 2
    R = 1 << (52 * s52)
    n = rnd.getrandbits(bits) | (1 << (bits - 1)) | 1
    np_ = (-pow(n, -1, 1 << 52)) & M52
    Rinv = pow(R, -1, n)
    cases = [(rnd.randrange(2 * n), rnd.randrange(2 * n)) for _ in range(2 if s52 > 80 else 4)]
    cases += [(0, 0), (2 * n - 1, 2 * n - 1), (1, n - 1), ((1 << bits) - 1, n - 1), (n, 1)]
    for a, b in cases:
        r = value(mont52_mul(limbs(a, s52), limbs(b, s52), limbs(n, s52), np_, tpi, L))
        assert r < 2 * n and r % n == a * b * Rinv % n
    # a chain of squarings keeps the lazy residue below 2n
    x = rnd.randrange(n)
    want = x
    for _ in range(3):
        x = value(mont52_mul(limbs(x, s52), limbs(x, s52), limbs(n, s52), np_, tpi, L))
        want = want * want * Rinv % n
        assert x < 2 * n and x % n == want
