"""GPU parity for the level-2 (mod n^3) path, alternative encryption, randomness extraction, nested
operations and the DDLEQ proofs against the Python oracle (oracle/paillier_ref.py), bit-exact.
Mirrors paillier_test.go:65-138, operations_test.go:56-163 and ddleq_test.go:9-72 with fixed randomness."""
import random

import pytest

from oracle import paillier_ref as R
from paillier_b200 import synth
from paillier_b200._lib import PgpuError, PGPU_ERR_ARG
from paillier_b200.api import (ALTERNATIVE, Ciphertext, DDLEQProof, DDLEQProofInstance, ENC_LEVEL_ONE, ENC_LEVEL_TWO,
                               PublicKey, SecretKey)

pytestmark = pytest.mark.gpu


def _units(rnd, n, k):
    out = []
    while len(out) < k:
        r = rnd.randrange(1, n)
        from math import gcd
        if gcd(r, n) == 1:
            out.append(r)
    return out


def _factors(n, lam):
    from math import isqrt
    s = n + 1 - lam                      # p + q
    d = isqrt(s * s - 4 * n)
    p, q = (s - d) // 2, (s + d) // 2
    assert p * q == n
    return p, q


def _keys(name, seed=7):
    p, q = synth.load_key(name)
    rnd = random.Random(seed)
    osk, opk = R.keygen_from_primes(p, q, h_seed_r=_units(rnd, p * q, 1)[0])
    sk = SecretKey(p * q, p=p, q=q, H=opk.H, K=opk.K)
    return sk, osk, opk, rnd


@pytest.fixture(scope="module", params=["paillier_64", "paillier_1024", "paillier_2048"])
def keys(request):
    sk, osk, opk, rnd = _keys(request.param)
    yield sk, osk, opk, rnd
    sk.close()


def test_level2_encrypt_decrypt(keys):
    sk, osk, opk, rnd = keys
    n, n2 = sk.N, sk.N ** 2
    ms = [0, 1, n2 - 1, n, n - 1] + [rnd.randrange(n2) for _ in range(6)]
    rs = [1, n - 1] + _units(rnd, n, len(ms) - 2)
    cts = sk.EncryptWithRAtLevelBatch(ms, rs, ENC_LEVEL_TWO)
    assert [c.C for c in cts] == [R.encrypt_with_r_at_level(opk, m, r, R.ENC_LEVEL_TWO).C for m, r in zip(ms, rs)]
    assert sk.DecryptBatch(cts) == ms == [R.decrypt(osk, R.Ciphertext(c.C, R.ENC_LEVEL_TWO)) for c in cts]
    # the public-key path (r^(n^2) mod n^3 directly) gives the same ciphertexts as the secret-key path over p^3, q^3
    assert [c.C for c in PublicKey.EncryptWithRAtLevelBatch(sk, ms, rs, ENC_LEVEL_TWO)] == [c.C for c in cts]
    # non-units and r around the prime powers: the reduced exponent must stay exact
    p, q = _factors(n, osk.Lambda)
    edge = [0, p, q, 2 * p, q * (p - 1), (p ** 3) % n, (q ** 3 + 1) % n, n // 2]
    em = [rnd.randrange(n2) for _ in edge]
    got = [c.C for c in sk.EncryptWithRAtLevelBatch(em, edge, ENC_LEVEL_TWO)]
    assert got == [R.encrypt_with_r_at_level(opk, m, r, R.ENC_LEVEL_TWO).C for m, r in zip(em, edge)]
    assert got == [c.C for c in PublicKey.EncryptWithRAtLevelBatch(sk, em, edge, ENC_LEVEL_TWO)]
    # level 1 through the same entry point
    ms1 = [m % n for m in ms]
    cts1 = sk.EncryptWithRAtLevelBatch(ms1, rs, ENC_LEVEL_ONE)
    assert [c.C for c in cts1] == [R.encrypt_with_r(opk, m, r).C for m, r in zip(ms1, rs)]


def test_nested_encrypt_decrypt(keys):
    # paillier_test.go:65-76,110-138: [[m]] = Enc2(Enc1(m)); NestedDecrypt peels both layers; inner 0 -> 0
    sk, osk, opk, rnd = keys
    n = sk.N
    ms = [5, 0, n - 1] + [rnd.randrange(n) for _ in range(3)]
    inner = sk.EncryptWithRBatch(ms, _units(rnd, n, len(ms)))
    outer = sk.EncryptWithRAtLevelBatch([c.C for c in inner] + [0], _units(rnd, n, len(ms) + 1), ENC_LEVEL_TWO)
    layer = sk.DecryptNestedCiphertextLayerBatch(outer)
    assert [c.C for c in layer] == [c.C for c in inner] + [0]
    assert sk.NestedDecryptBatch(outer) == ms + [0]
    assert sk.NestedDecryptBatch(outer) == [R.nested_decrypt(osk, R.Ciphertext(c.C, R.ENC_LEVEL_TWO)) for c in outer]
    with pytest.raises(ValueError):
        sk.DecryptNestedCiphertextLayerBatch(inner)


def test_alt_encrypt(keys):
    sk, osk, opk, rnd = keys
    n, n2 = sk.N, sk.N ** 2
    for level, space in ((ENC_LEVEL_ONE, n), (ENC_LEVEL_TWO, n2)):
        ms = [0, 1, space - 1] + [rnd.randrange(space) for _ in range(5)]
        rs = [0, 1, opk.K, opk.K - 1] + [rnd.randrange(n) for _ in range(4)]
        mine = list(rs)
        cts = sk.AltEncryptWithRAtLevelBatch(ms, mine, level)
        ref = [R.alt_encrypt_with_r_at_level(opk, m, r, level) for m, r in zip(ms, rs)]
        assert [c.C for c in cts] == [c.C for c, _ in ref]
        assert mine == [r for _, r in ref]                      # r reduced mod K in place (paillier.go:228)
        assert all(c.EncMethod == ALTERNATIVE for c in cts)
        assert sk.DecryptBatch(cts) == ms


def test_randomize_and_extract_randomness(keys):
    # operations_test.go:130-163: ExtractRandonness returns the r of EncryptWithRAtLevel
    sk, osk, opk, rnd = keys
    n, n2 = sk.N, sk.N ** 2
    ms = [0, 1, n - 1] + [rnd.randrange(n) for _ in range(4)]
    rs = [1, 4, 9] + _units(rnd, n, 4)
    cts = sk.EncryptWithRBatch(ms, rs)
    assert sk.ExtractRandonnessBatch(cts) == rs == [R.extract_randomness(osk, R.Ciphertext(c.C)) for c in cts]
    ms2 = [m * 3 % n2 for m in ms]
    cts2 = sk.EncryptWithRAtLevelBatch(ms2, rs, ENC_LEVEL_TWO)
    assert sk.ExtractRandonnessBatch(cts2) == rs == [R.extract_randomness(osk, R.Ciphertext(c.C, R.ENC_LEVEL_TWO)) for c in cts2]
    r2 = _units(rnd, n, len(cts))
    rz = sk.RandomizeWithRBatch(cts, r2)
    assert [c.C for c in rz] == [R.add(opk, R.Ciphertext(c.C), R.encrypt_with_r(opk, 0, r)).C for c, r in zip(cts, r2)]
    assert sk.DecryptBatch(rz) == ms
    assert sk.ExtractRandonnessBatch(rz) == [a * b % n for a, b in zip(rs, r2)]
    pk = PublicKey(n)                      # public-key route (r^n mod n^2 directly) against the key holder's (over p^2, q^2)
    assert [c.C for c in pk.RandomizeWithRBatch(cts, r2)] == [c.C for c in rz]
    pk.close()


def test_nested_operations(keys):
    # operations_test.go:56-128
    sk, osk, opk, rnd = keys
    n = sk.N
    ms = [rnd.randrange(n // 4) for _ in range(4)]
    ks = [rnd.randrange(n // 4) for _ in range(4)]
    inner = sk.EncryptWithRBatch(ms, _units(rnd, n, 4))
    outer = sk.EncryptWithRAtLevelBatch([c.C for c in inner], _units(rnd, n, 4), ENC_LEVEL_TWO)
    addend = sk.EncryptWithRBatch(ks, _units(rnd, n, 4))
    o2 = lambda c: R.Ciphertext(c.C, R.ENC_LEVEL_TWO)
    o1 = lambda c: R.Ciphertext(c.C, R.ENC_LEVEL_ONE)
    added = sk.NestedAddBatch(outer, addend)
    assert [c.C for c in added] == [R.nested_add(opk, o2(a), o1(b)).C for a, b in zip(outer, addend)]
    assert sk.NestedDecryptBatch(added) == [(m + k) % n for m, k in zip(ms, ks)]
    subbed = sk.NestedSubBatch(outer, addend)
    assert [c.C for c in subbed] == [R.nested_sub(opk, o2(a), o1(b)).C for a, b in zip(outer, addend)]
    assert sk.NestedDecryptBatch(subbed) == [(m - k) % n for m, k in zip(ms, ks)]
    As, Bs = _units(rnd, n, 4), _units(rnd, n, 4)
    rz = sk.NestedRandomizeWithBatch(outer, As, Bs)
    assert [c.C for c in rz] == [R.nested_randomize_with(opk, o2(c), a, b).C for c, a, b in zip(outer, As, Bs)]
    assert sk.NestedDecryptBatch(rz) == ms
    with pytest.raises(ValueError):
        sk.NestedRandomizeWithBatch(inner, As, Bs)
    with pytest.raises(ValueError):
        sk.NestedAddBatch(inner, addend)


def test_ddleq_prove_verify(keys):
    # ddleq_test.go:9-72 with the instance randomness fixed: transcripts bit-exact with the oracle
    sk, osk, opk, rnd = keys
    n = sk.N
    big = n.bit_length() > 1500
    count, secpar = (2, 4) if big else (3, 10)
    ms = [rnd.randrange(n) for _ in range(count)]
    inner = sk.EncryptWithRBatch(ms, _units(rnd, n, count))
    ct1 = sk.EncryptWithRAtLevelBatch([c.C for c in inner], _units(rnd, n, count), ENC_LEVEL_TWO)
    As, Bs = _units(rnd, n, count), _units(rnd, n, count)
    ct2 = sk.NestedRandomizeWithBatch(ct1, As, Bs)
    xs = [_units(rnd, n, secpar) for _ in range(count)]
    ys = [_units(rnd, n, secpar) for _ in range(count)]
    proofs = sk.ProveDDLEQBatch(secpar, ct1, ct2, As, Bs, xs, ys)
    o2 = lambda c: R.Ciphertext(c.C, R.ENC_LEVEL_TWO)
    chal = set()
    for i in range(count if not big else 1):
        ref = R.prove_ddleq(osk, secpar, o2(ct1[i]), o2(ct2[i]), As[i], Bs[i], xs[i], ys[i])
        for g, o in zip(proofs[i].Instances, ref):
            assert (g.X, g.Y, g.Alpha, g.E, g.F) == (o.X, o.Y, o.Alpha, o.E, o.F)
            chal.add(o.E != o.X)
        assert R.verify_ddleq(opk, o2(ct1[i]), o2(ct2[i]), ref)
    if not big:
        assert chal == {True, False}          # both challenge values exercised
    assert sk.VerifyDDLEQProofBatch(ct1, ct2, proofs) == [True] * count
    # the prover works over p^3, q^3 (CRT); the direct route over n^3 must give the same transcripts
    import os
    os.environ["PGPU_NO_CRT_PROTOCOLS"] = "1"
    try:
        direct = sk.ProveDDLEQBatch(secpar, ct1, ct2, As, Bs, xs, ys)
    finally:
        del os.environ["PGPU_NO_CRT_PROTOCOLS"]
    assert direct == proofs
    # soundness (ddleq_test.go:38-72): a proof for other ciphertexts / a tampered instance must be rejected
    bad = [DDLEQProof(list(p.Instances)) for p in proofs]
    i0 = bad[0].Instances[1]
    bad[0].Instances[1] = DDLEQProofInstance(i0.X, i0.Y, i0.Alpha, i0.E, (i0.F + 1) % n ** 3)
    assert sk.VerifyDDLEQProofBatch(ct1, ct2, bad) == [False] + [True] * (count - 1)
    swapped = sk.VerifyDDLEQProofBatch(ct2, ct1, proofs)
    assert swapped == [False] * count
    # wrong (a, b): the reference panics (ddleq.go:67-69)
    with pytest.raises(PgpuError) as ei:
        sk.ProveDDLEQBatch(secpar, ct1, ct2, As[::-1], Bs, xs, ys)
    assert ei.value.code == PGPU_ERR_ARG and "inputs are wrong" in str(ei.value)


def test_nested_randomize_three_routes(keys):
    """NestedRandomize (operations.go:96-118) = ct^(a^n mod n^2) * b^(n^2) mod n^3: the key holder's route over p^3, q^3,
    the public-key route with interleaved (shared-squaring) exponentiation and the plain route give the same bits;
    DDLEQ verification accepts/rejects identically with and without the interleaved exponentiation."""
    import os
    sk, osk, opk, rnd = keys
    n = sk.N
    n2, n3 = n * n, n ** 3
    count = 9
    cts = [Ciphertext(rnd.randrange(1, n3), ENC_LEVEL_TWO) for _ in range(count)]
    As, Bs = _units(rnd, n, count), _units(rnd, n, count)
    want = [pow(c.C, pow(a, n, n2), n3) * pow(b, n2, n3) % n3 for c, a, b in zip(cts, As, Bs)]
    pk = PublicKey(n)
    try:
        assert [c.C for c in sk.NestedRandomizeWithBatch(cts, As, Bs)] == want          # CRT
        assert [c.C for c in pk.NestedRandomizeWithBatch(cts, As, Bs)] == want          # interleaved
        os.environ["PGPU_NO_DUAL_EXP"] = "1"
        os.environ["PGPU_NO_CRT_PROTOCOLS"] = "1"
        assert [c.C for c in pk.NestedRandomizeWithBatch(cts, As, Bs)] == want          # plain
        assert [c.C for c in sk.NestedRandomizeWithBatch(cts, As, Bs)] == want
        del os.environ["PGPU_NO_DUAL_EXP"], os.environ["PGPU_NO_CRT_PROTOCOLS"]
        # a small DDLEQ statement verified through both routes, honest and tampered
        inner = sk.EncryptWithRBatch([5], _units(rnd, n, 1))
        ct1 = sk.EncryptWithRAtLevelBatch([inner[0].C], _units(rnd, n, 1), ENC_LEVEL_TWO)
        a, b = _units(rnd, n, 1), _units(rnd, n, 1)
        ct2 = sk.NestedRandomizeWithBatch(ct1, a, b)
        proofs = sk.ProveDDLEQBatch(6, ct1, ct2, a, b, [_units(rnd, n, 6)], [_units(rnd, n, 6)])
        bad = [DDLEQProof(list(proofs[0].Instances))]
        i0 = bad[0].Instances[2]
        bad[0].Instances[2] = DDLEQProofInstance(i0.X, i0.Y, i0.Alpha, (i0.E + 1) % n2, i0.F)
        for env in (None, "1"):
            if env:
                os.environ["PGPU_NO_DUAL_EXP"] = env
            assert pk.VerifyDDLEQProofBatch(ct1, ct2, proofs) == [True]
            assert pk.VerifyDDLEQProofBatch(ct1, ct2, bad) == [False]
    finally:
        os.environ.pop("PGPU_NO_DUAL_EXP", None)
        os.environ.pop("PGPU_NO_CRT_PROTOCOLS", None)
        pk.close()


def test_homomorphic_ops_follow_the_ciphertext_level(keys):
    # operations.go:11-64 take n^(s+1) from the level of the (first) ciphertext: Add / Sub / ConstMult at level 2 work
    # mod n^3 and keep the level (the mirrors used n^2 regardless)
    sk, osk, opk, rnd = keys
    n, n2, n3 = sk.N, sk.N ** 2, sk.N ** 3
    count = 7
    ms = [rnd.randrange(n2) for _ in range(count)]
    ks = [0, 1, 2, rnd.randrange(1 << 64), rnd.randrange(1 << 200), n - 1, 3]
    cts = sk.EncryptWithRAtLevelBatch(ms, _units(rnd, n, count), ENC_LEVEL_TWO)
    other = sk.EncryptWithRAtLevelBatch(ms[::-1], _units(rnd, n, count), ENC_LEVEL_TWO)
    o2 = lambda c: R.Ciphertext(c.C, R.ENC_LEVEL_TWO, c.EncMethod)
    cm = sk.ConstMultBatch(cts, ks)
    assert [(c.C, c.Level) for c in cm] == [(R.const_mult(opk, o2(c), k).C, ENC_LEVEL_TWO) for c, k in zip(cts, ks)]
    assert sk.DecryptBatch(cm[1:]) == [m * k % n2 for m, k in zip(ms[1:], ks[1:])]
    tot = sk.AddBatch(cts)
    assert (tot.C, tot.Level) == (R.add(opk, *[o2(c) for c in cts]).C, ENC_LEVEL_TWO)
    assert sk.DecryptBatch([tot]) == [sum(ms) % n2]
    ap = sk.AddPairs(cts, other)
    assert [(c.C, c.Level) for c in ap] == [(R.add(opk, o2(a), o2(b)).C, ENC_LEVEL_TWO) for a, b in zip(cts, other)]
    sp = sk.SubPairs(cts, other)
    assert [(c.C, c.Level) for c in sp] == [(R.sub(opk, o2(a), o2(b)).C, ENC_LEVEL_TWO) for a, b in zip(cts, other)]
    assert sk.DecryptBatch(sp) == [(a - b) % n2 for a, b in zip(ms, ms[::-1])]
    # a level-1 batch still works mod n^2, and a batch mixing levels is refused
    l1 = sk.EncryptWithRBatch([5, 6], _units(rnd, n, 2))
    assert sk.AddBatch(l1).C == R.add(opk, *[R.Ciphertext(c.C) for c in l1]).C
    with pytest.raises(ValueError):
        sk.ConstMultBatch([cts[0], l1[0]], [1, 2])
    with pytest.raises(ValueError):
        sk.AddPairs([cts[0], l1[0]], [cts[1], l1[1]])
    # Add takes the level of its first argument (operations.go:13): a level-1 value multiplied in mod n^3
    mixed = sk.AddBatch([cts[0], l1[0]])
    assert (mixed.C, mixed.Level) == (R.add(opk, o2(cts[0]), R.Ciphertext(l1[0].C)).C, ENC_LEVEL_TWO)
