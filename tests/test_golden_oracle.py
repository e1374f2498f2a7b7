"""CPU: the two oracles against the committed vectors (tests/golden/vectors.json, made by tools/gen_golden.py).
Python oracle and libgmp oracle are unrelated bignum implementations; agreeing on these pins both."""
import json
import os

import numpy as np
import pytest

from oracle import gmp_ref as G
from oracle import paillier_ref as R

V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
I = lambda s: int(s, 16)


def _rec(vals, w):
    return np.frombuffer(b"".join(int(v).to_bytes(w, "little") for v in vals), dtype=np.uint8).copy()


def _unrec(buf, w):
    b = bytes(buf)
    return [int.from_bytes(b[i:i + w], "little") for i in range(0, len(b), w)]


@pytest.mark.parametrize("name", ["paillier_64", "paillier_1024", "paillier_2048"])
def test_encrypt_vectors_both_oracles(name):
    c = V["cases"][name]
    p, q = I(c["p"]), I(c["q"])
    n = p * q
    sk, pk = R.keygen_from_primes(p, q)
    ms, rs, cs = [I(x) for x in c["encrypt"]["m"]], [I(x) for x in c["encrypt"]["r"]], [I(x) for x in c["encrypt"]["c"]]
    assert [R.encrypt_with_r(pk, m, r).C for m, r in zip(ms, rs)] == cs
    assert [R.decrypt(sk, R.Ciphertext(x)) for x in cs] == ms
    w = max(8, (n.bit_length() + 63) // 64 * 8)
    got = G.encrypt_with_r(n, _rec(ms, w), _rec(rs, w), w)
    assert _unrec(got, 2 * w) == cs
    assert _unrec(G.decrypt(n, sk.Lambda, got, w), w) == ms
    ks = [I(x) for x in c["const_mult"]["k"]]
    assert _unrec(G.modexp(n * n, _rec(cs[:4], 2 * w), 2 * w, _rec(ks, 8), 8), 2 * w) == [I(x) for x in c["const_mult"]["c"]]
    assert _unrec(G.add_reduce(n * n, _rec(cs, 2 * w), 2 * w), 2 * w) == [I(c["add_all"])]


def test_level2_alt_ddleq_vectors_small_key():
    c = V["cases"]["paillier_64"]
    p, q = I(c["p"]), I(c["q"])
    sk, pk = R.keygen_from_primes(p, q)
    pk.H = sk.H = I(c["H"])
    m2, rs = [I(x) for x in c["encrypt_level2"]["m"]], [I(x) for x in c["encrypt_level2"]["r"]]
    assert [R.encrypt_with_r_at_level(pk, m, r, R.ENC_LEVEL_TWO).C for m, r in zip(m2, rs)] == [I(x) for x in c["encrypt_level2"]["c"]]
    ms = [I(x) for x in c["encrypt"]["m"]]
    ar = [I(x) for x in c["alt_encrypt"]["r"]]
    assert [R.alt_encrypt_with_r_at_level(pk, m, r, R.ENC_LEVEL_ONE)[0].C for m, r in zip(ms, ar)] == [I(x) for x in c["alt_encrypt"]["c_level1"]]
    d = c["ddleq"]
    proof = R.prove_ddleq(sk, 4, R.Ciphertext(I(d["ct1"]), R.ENC_LEVEL_TWO), R.Ciphertext(I(d["ct2"]), R.ENC_LEVEL_TWO), I(d["a"]), I(d["b"]),
                          [I(x) for x in d["x"]], [I(x) for x in d["y"]])
    assert [(i.Alpha, i.E, i.F) for i in proof] == list(zip(*[[I(x) for x in d[k]] for k in ("alpha", "e", "f")]))


def test_threshold_vectors_small_key():
    c = V["cases"]["threshold_512"]
    p, q = I(c["p"]), I(c["q"])
    n = p * q
    key = R.ThresholdSecretKey(N=n, TotalNumberOfDecryptionServers=c["l"], Threshold=c["w"], VerificationKey=I(c["V"]),
                               VerificationKeys=[I(x) for x in c["vi"]], ID=2, Share=I(c["shares"][1]))
    for cc, r, dec, e, z in zip(c["c"], c["zkp_r"], c["partial_decrypt_id2"], c["zkp_e"], c["zkp_z"]):
        zk = R.partial_decryption_with_zkp(key, I(cc), I(r))
        assert (zk.Decryption, zk.E, zk.Z) == (I(dec), I(e), I(z))
    w2 = 128
    got = G.partial_decrypt(n, key.Share, c["l"], _rec([I(x) for x in c["c"]], w2), w2)
    assert _unrec(got, w2) == [I(x) for x in c["partial_decrypt_id2"]]


def test_safe_prime_vectors():
    for bits, d in V["safe_prime"].items():
        if int(bits) > 64:
            continue
        for raw, p, q, ok in zip(d["raw"], d["p"], d["q"], d["ok"]):
            assert R.safe_prime_candidate(bytes.fromhex(raw), int(bits)) == (I(p), I(q), ok)
    assert sum(V["safe_prime"]["64"]["ok"]) >= 2
