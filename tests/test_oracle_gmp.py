"""The libgmp restatement (oracle/gmp_ref.c, also the CPU baseline) against the Python oracle."""
import numpy as np

from oracle import gmp_ref as G
from oracle import paillier_ref as R
from paillier_b200 import synth
from paillier_b200.api import from_records, to_records


def _key(name):
    p, q = synth.load_key(name)
    return p, q, p * q


def test_encrypt_decrypt_calls_match_python_oracle():
    for name in ("paillier_64", "paillier_1024", "paillier_2048"):
        p, q, n = _key(name)
        sk, pk = R.keygen_from_primes(p, q)
        w_n = {64: 64, 1024: 128, 2048: 256}[n.bit_length()]
        count = 6
        m = synth.plaintexts(count, n, w_n)
        r = synth.randomness(count, n, w_n)
        ms, rs = from_records(m, w_n), from_records(r, w_n)
        ms[0], ms[1], rs[2] = 0, n - 1, 1
        m, r = to_records(ms, w_n), to_records(rs, w_n)
        c = G.encrypt_with_r(n, m, r, w_n, threads=2)
        exp = [R.encrypt_with_r(pk, mi, ri).C for mi, ri in zip(ms, rs)]
        assert from_records(c, 2 * w_n) == exp
        d = G.decrypt(n, sk.Lambda, c, w_n, threads=3)
        assert from_records(d, w_n) == ms == [R.decrypt(sk, R.Ciphertext(ci)) for ci in exp]


def test_partial_decrypt_kat_and_modexp():
    # thresholdkey_test.go:58-74 through the libgmp path
    out = G.partial_decrypt(101 * 103, 862, 10, to_records([56], 128), 128, threads=1)
    assert from_records(out, 128) == [40644522]
    p, q, n = _key("paillier_1024")
    n2 = n * n
    bases = [pow(3, i + 1, n2) for i in range(5)]
    exps = [0, 1, 2 ** 64 - 1, 12345678901234567890, 2 ** 63]
    out = G.modexp(n2, to_records(bases, 256), 256, to_records(exps, 8), 8, threads=2)
    assert from_records(out, 256) == [R.gmp_exp(b, e, n2) for b, e in zip(bases, exps)]
    out = G.modmul(n2, to_records(bases, 256), to_records(bases[::-1], 256), 256)
    assert from_records(out, 256) == [a * b % n2 for a, b in zip(bases, bases[::-1])]
    out = G.add_reduce(n2, to_records(bases, 256), 256, threads=3)
    pk = R.PublicKey(N=n)
    assert from_records(out, 256) == [R.add(pk, *[R.Ciphertext(b) for b in bases]).C]


def test_modinv_and_sub_call_sequence_match_python_oracle():
    # Sub (operations.go:32-55): ct1 * ModInverse(ct2, n^2) through mpz_invert / mpz_mul / mpz_mod at 2048 bits
    p, q, n = _key("paillier_2048")
    n2 = n * n
    pk = R.PublicKey(N=n)
    a = [pow(5, i + 3, n2) for i in range(6)]
    b = [pow(7, 2 * i + 1, n2) for i in range(6)]
    inv, ok = G.modinv(n2, to_records(b, 512), 512, threads=2)
    assert ok.tolist() == [1] * 6
    assert from_records(inv, 512) == [pow(x, -1, n2) for x in b]
    out = G.modmul(n2, to_records(a, 512), inv, 512)
    assert from_records(out, 512) == [R.sub(pk, R.Ciphertext(x), R.Ciphertext(y)).C for x, y in zip(a, b)]
    # non-units are reported, not inverted (mpz_invert returns 0)
    inv, ok = G.modinv(n2, to_records([p, 2, q * q, 1], 512), 512, threads=1)
    assert ok.tolist() == [0, 1, 0, 1]
    assert from_records(inv, 512) == [0, pow(2, -1, n2), 0, 1]
