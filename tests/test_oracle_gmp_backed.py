"""The Python oracle (oracle/paillier_ref.py) funnels every Exp / ModInverse through gmp_exp / gmp_mod_inverse.  Here the
same restatement runs a second time with those two primitives answered by libgmp (oracle/gmp_ref.c: mpz_powm, mpz_invert)
instead of CPython's pow, at 2048 / 3072-bit key sizes and fixed randomness, for every row that the call-sequence functions
of gmp_ref.c do not cover: level 2, alternative encryption, randomness extraction, nested operations, the threshold ZKP,
share combining and the DDLEQ transcripts.  Two unrelated bignum implementations must produce the same transcripts."""
import random
from math import gcd

import pytest

from oracle import gmp_ref as G
from oracle import paillier_ref as R
from paillier_b200 import synth
from paillier_b200.api import from_records, to_records

CALLS = {"exp": 0, "inv": 0}


def _libgmp_exp(x, y, m):
    if y <= 0:
        return 1                                   # ncw/gmp Int.Exp, as in R.gmp_exp
    if m is None or m == 0:
        return x ** y
    m = abs(m)
    w = (m.bit_length() + 63) // 64 * 8
    eb = max(8, (y.bit_length() + 63) // 64 * 8)
    CALLS["exp"] += 1
    return from_records(G.modexp(m, to_records([x % m], w), w, to_records([y], eb), eb, threads=1), w)[0]


def _libgmp_inv(a, m):
    w = (m.bit_length() + 63) // 64 * 8
    out, ok = G.modinv(m, to_records([a % m], w), w, threads=1)
    if not ok[0]:
        raise ValueError("base is not invertible for the given modulus")
    CALLS["inv"] += 1
    return from_records(out, w)[0]


def _units(rnd, n, k):
    out = []
    while len(out) < k:
        r = rnd.randrange(1, n)
        if gcd(r, n) == 1:
            out.append(r)
    return out


def _scheme_transcript(p, q, seed):
    rnd = random.Random(seed)
    n = p * q
    sk, pk = R.keygen_from_primes(p, q, h_seed_r=_units(rnd, n, 1)[0])
    n2 = n * n
    out = []
    m1, m2 = rnd.randrange(n), rnd.randrange(n2)
    r1, r2, ra = _units(rnd, n, 1)[0], _units(rnd, n, 1)[0], rnd.randrange(n)
    c1 = R.encrypt_with_r(pk, m1, r1)
    c2 = R.encrypt_with_r_at_level(pk, m2, r2, R.ENC_LEVEL_TWO)
    out += [c1.C, c2.C, R.decrypt(sk, c1), R.decrypt(sk, c2)]
    for level, m in ((R.ENC_LEVEL_ONE, m1), (R.ENC_LEVEL_TWO, m2)):
        ct, rr = R.alt_encrypt_with_r_at_level(pk, m, ra, level)
        out += [ct.C, rr, R.decrypt(sk, ct)]
    out += [R.extract_randomness(sk, c1), R.extract_randomness(sk, c2)]
    out += [R.sub(pk, c1, R.encrypt_with_r(pk, 5, r2)).C, R.const_mult(pk, c1, 2 ** 64 - 3).C]
    outer = R.encrypt_with_r_at_level(pk, c1.C, r2, R.ENC_LEVEL_TWO)
    a, b = _units(rnd, n, 2)
    rerand = R.nested_randomize_with(pk, outer, a, b)
    out += [outer.C, rerand.C, R.nested_decrypt(sk, rerand)]
    out += [R.nested_add(pk, outer, c1).C, R.nested_sub(pk, outer, c1).C]
    xs, ys = _units(rnd, n, 3), _units(rnd, n, 3)
    proof = R.prove_ddleq(sk, 3, outer, rerand, a, b, xs, ys)
    for inst in proof:
        out += [inst.Alpha, inst.E, inst.F]
    out.append(int(R.verify_ddleq(pk, outer, rerand, proof)))
    return out


def _threshold_transcript(p, q, seed):
    rnd = random.Random(seed)
    n = p * q
    nm = n * ((p - 1) // 2) * ((q - 1) // 2)
    keys = R.threshold_keys_from(p, q, 4, 3, v_seed=rnd.randrange(2, n * n), coeffs=[rnd.randrange(nm) for _ in range(2)])
    pk = R.PublicKey(N=n)
    m = rnd.randrange(n)
    c = R.encrypt_with_r(pk, m, _units(rnd, n, 1)[0]).C
    out = [keys[0].VerificationKey] + list(keys[0].VerificationKeys)
    zk = [R.partial_decryption_with_zkp(k, c, rnd.randrange(n * n)) for k in keys[:3]]
    for z in zk:
        out += [z.Decryption, z.E, z.Z, int(R.verify_proof(z))]
    tk = R.threshold_public_key(keys[0])
    out.append(R.combine_partial_decryptions_zkp(tk, zk))
    assert out[-1] == m
    return out


@pytest.mark.parametrize("which,key", [("scheme", "paillier_2048"), ("scheme", "paillier_1024"), ("threshold", "threshold_2048"),
                                       ("threshold", "threshold_3072")])
def test_transcripts_identical_under_libgmp_primitives(monkeypatch, which, key):
    p, q = synth.load_key(key)
    fn = _scheme_transcript if which == "scheme" else _threshold_transcript
    with_cpython = fn(p, q, seed=17)
    monkeypatch.setattr(R, "gmp_exp", _libgmp_exp)
    monkeypatch.setattr(R, "gmp_mod_inverse", _libgmp_inv)
    before = dict(CALLS)
    with_libgmp = fn(p, q, seed=17)
    assert CALLS["exp"] > before["exp"] + 10, "the libgmp primitives were not reached"
    assert with_libgmp == with_cpython
