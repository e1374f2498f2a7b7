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


def _threshold_key(name, l, w, seed):
    import random
    p, q, n = _key(name)
    rnd = random.Random(seed)
    nm = n * ((p - 1) // 2) * ((q - 1) // 2)
    keys = R.threshold_keys_from(p, q, l, w, v_seed=rnd.randrange(2, n * n), coeffs=[rnd.randrange(nm) for _ in range(w - 1)])
    return n, keys, rnd


def test_zkp_call_sequences_match_python_oracle():
    # PartialDecryptionWithZKP / VerifyProof (thresholdkey.go:225-326) through libgmp + libcrypto's SHA-256
    for name, l, w, w2 in (("threshold_512", 5, 3, 128), ("threshold_2048", 8, 5, 512)):
        n, keys, rnd = _threshold_key(name, l, w, 11)
        pk = R.PublicKey(N=n)
        k = keys[2]
        wz = w2 + 64
        cs = [R.encrypt_with_r(pk, rnd.randrange(n), rnd.randrange(1, n)).C for _ in range(4)]
        rs = [0] + [rnd.randrange(n * n) for _ in range(3)]
        dec, e, z = G.pdec_zkp(n, k.Share, l, k.VerificationKey, to_records(cs, w2), to_records(rs, w2), w2, wz, threads=2)
        want = [R.partial_decryption_with_zkp(k, c, r) for c, r in zip(cs, rs)]
        assert from_records(dec, w2) == [x.Decryption for x in want]
        assert from_records(e, 32) == [x.E for x in want]
        assert from_records(z, wz) == [x.Z for x in want]
        vi = k.VerificationKeys[k.ID - 1]
        ok = G.zkp_verify(n, k.VerificationKey, vi, to_records(cs, w2), dec, e, z, w2, wz, threads=2)
        assert ok.tolist() == [1] * 4 == [int(R.verify_proof(x)) for x in want]
        bad_e = e.copy(); bad_e[32] ^= 1                                   # proof 1: wrong E
        bad_dec = dec.copy(); bad_dec[2 * w2] ^= 1                          # proof 2: wrong partial decryption
        assert G.zkp_verify(n, k.VerificationKey, vi, to_records(cs, w2), bad_dec, bad_e, z, w2, wz).tolist() == [1, 0, 0, 1]
        other = k.VerificationKeys[k.ID % l]                                # another server's verification key
        assert G.zkp_verify(n, k.VerificationKey, other, to_records(cs, w2), dec, e, z, w2, wz).tolist() == [0] * 4


def test_ddleq_verify_dot_and_safe_prime_call_sequences():
    import random
    from math import gcd
    p, q, n = _key("paillier_1024")
    n2 = n * n
    rnd = random.Random(3)
    sk, pk = R.keygen_from_primes(p, q)
    units = lambda k: [x for x in (rnd.randrange(1, n) for _ in range(4 * k)) if gcd(x, n) == 1][:k]
    secpar = 6
    r1, r2, a, b = units(4)
    ct1 = R.encrypt_with_r_at_level(pk, R.encrypt_with_r(pk, rnd.randrange(n), r1).C, r2, R.ENC_LEVEL_TWO)
    ct2 = R.nested_randomize_with(pk, ct1, a, b)
    xs, ys = units(secpar), units(secpar)
    proof = R.prove_ddleq(sk, secpar, ct1, ct2, a, b, xs, ys)
    wn, w2, w3 = 128, 256, 384
    args = (to_records([ct1.C], w3), to_records([ct2.C], w3), to_records(xs, wn), to_records(ys, wn),
            to_records([i.Alpha for i in proof], w3), to_records([i.E for i in proof], w2), to_records([i.F for i in proof], w3))
    assert G.ddleq_verify(n, secpar, *args, wn, w2, w3, threads=2).tolist() == [1] * secpar
    assert len({R.random_oracle_bit(ct1.C, ct2.C, i.X, i.Y, i.Alpha) for i in proof}) == 2      # both challenge values
    f_bad = to_records([proof[0].F ^ 1] + [i.F for i in proof[1:]], w3)
    assert G.ddleq_verify(n, secpar, *args[:6], f_bad, wn, w2, w3).tolist() == [0] + [1] * (secpar - 1)
    # encrypted dot product = ConstMult per term folded by Add
    cs = [R.encrypt_with_r(pk, rnd.randrange(n), r).C for r in units(9)]
    ks = [0, 1, 2 ** 64 - 1] + [rnd.getrandbits(64) for _ in range(6)]
    want = R.add(pk, *[R.const_mult(pk, R.Ciphertext(c), k) for c, k in zip(cs, ks)]).C
    assert from_records(G.dot_u64(n2, to_records(cs, w2), w2, np.array(ks, dtype=np.uint64), threads=3), w2) == [want]
    # safe-prime candidate procedure: same accept / reject decisions as the Python oracle on the committed byte strings
    import json, os
    V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))["safe_prime"]
    for bits in ("16", "64", "1024"):
        raw = b"".join(bytes.fromhex(x) for x in V[bits]["raw"])
        assert G.safe_prime_scan(int(bits), raw, threads=2).tolist() == [int(x) for x in V[bits]["ok"]]
