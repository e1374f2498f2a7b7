"""GPU: the engine against the committed vectors (tests/golden/vectors.json) through the C ABI.  No oracle is imported here."""
import json
import os

import pytest

from paillier_b200.api import (Ciphertext, DDLEQProof, DDLEQProofInstance, ENC_LEVEL_ONE, ENC_LEVEL_TWO, PartialDecryption, SecretKey,
                               ThresholdSecretKey)
from paillier_b200.keygen import safe_prime_scan

pytestmark = pytest.mark.gpu
V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
I = lambda s: int(s, 16)
L = lambda xs: [I(x) for x in xs]


@pytest.mark.parametrize("name", ["paillier_64", "paillier_1024", "paillier_2048"])
def test_paillier_vectors(name):
    c = V["cases"][name]
    p, q = I(c["p"]), I(c["q"])
    sk = SecretKey(p * q, p=p, q=q, H=I(c["H"]), K=I(c["K"]))
    ms, rs, cs = L(c["encrypt"]["m"]), L(c["encrypt"]["r"]), L(c["encrypt"]["c"])
    cts = sk.EncryptWithRBatch(ms, rs)
    assert [x.C for x in cts] == cs and sk.DecryptBatch(cts) == ms
    m2, r2, c2 = L(c["encrypt_level2"]["m"]), L(c["encrypt_level2"]["r"]), L(c["encrypt_level2"]["c"])
    ct2 = sk.EncryptWithRAtLevelBatch(m2, r2, ENC_LEVEL_TWO)
    assert [x.C for x in ct2] == c2 and sk.DecryptBatch(ct2) == m2
    ar = L(c["alt_encrypt"]["r"])
    assert [x.C for x in sk.AltEncryptWithRAtLevelBatch(ms[:3], list(ar), ENC_LEVEL_ONE)] == L(c["alt_encrypt"]["c_level1"])
    assert [x.C for x in sk.AltEncryptWithRAtLevelBatch(m2[:3], list(ar), ENC_LEVEL_TWO)] == L(c["alt_encrypt"]["c_level2"])
    assert [x.C for x in sk.ConstMultBatch(cts[:4], L(c["const_mult"]["k"]))] == L(c["const_mult"]["c"])
    assert sk.AddBatch(cts).C == I(c["add_all"])
    assert [x.C for x in sk.SubPairs(cts[:3], cts[3:])] == L(c["sub_pairs"])
    assert sk.ExtractRandonnessBatch(ct2[:2]) == L(c["extract_randomness_level2"]) == r2[:2]
    d = c["ddleq"]
    ct1, ctb = Ciphertext(I(d["ct1"]), ENC_LEVEL_TWO), Ciphertext(I(d["ct2"]), ENC_LEVEL_TWO)
    proof = sk.ProveDDLEQBatch(4, [ct1], [ctb], [I(d["a"])], [I(d["b"])], [L(d["x"])], [L(d["y"])])[0]
    assert [(i.Alpha, i.E, i.F) for i in proof.Instances] == list(zip(L(d["alpha"]), L(d["e"]), L(d["f"])))
    assert sk.VerifyDDLEQProofBatch([ct1], [ctb], [proof]) == [True]
    sk.close()


@pytest.mark.parametrize("name", ["threshold_512", "threshold_2048", "threshold_3072"])
def test_threshold_vectors(name):
    c = V["cases"][name]
    n = I(c["p"]) * I(c["q"])
    shares = L(c["shares"])
    keys = [ThresholdSecretKey(n, c["l"], c["w"], I(c["V"]), L(c["vi"]), ID=i + 1, Share=shares[i]) for i in range(c["w"])]
    cs = L(c["c"])
    zk = keys[1].PartialDecryptionWithZKPBatch(cs, L(c["zkp_r"]))
    assert [z.Decryption for z in zk] == L(c["partial_decrypt_id2"])
    assert [z.E for z in zk] == L(c["zkp_e"]) and [z.Z for z in zk] == L(c["zkp_z"])
    assert all(keys[0].VerifyProofBatch(zk))
    parts = [k.PartialDecryptBatch(cs) for k in keys]
    assert keys[0].CombinePartialDecryptionsBatch(parts) == L(c["m"])
    for k in keys:
        k.close()


def test_safe_prime_vectors():
    for bits, d in V["safe_prime"].items():
        ps, qs, ok = safe_prime_scan(int(bits), b"".join(bytes.fromhex(r) for r in d["raw"]))
        assert qs == L(d["q"]) and ok == d["ok"]
        assert [p for p, o in zip(ps, L(d["p"])) if o] == [o for o in L(d["p"]) if o]
