"""The reference's convenience callers around the hot path -- the methods that draw their own randomness or take a variadic
argument list (paillier.go:192-203,244-289, operations.go:32-55,67-69,96-118, thresholdkey.go:213-222,258-275) -- through
the batch mirror, checked against the oracle (paillier_test.go:65-138, operations_test.go:30-128)."""
import random

import pytest

from oracle import paillier_ref as R
from paillier_b200._lib import PgpuError, PGPU_ERR_NOT_INVERTIBLE
from paillier_b200.api import ALTERNATIVE, Ciphertext, ENC_LEVEL_ONE, ENC_LEVEL_TWO, MIXED, REGULAR
from paillier_b200.keygen import KeyGen, ThresholdKeyGenerator

pytestmark = pytest.mark.gpu


def test_drawing_encryptions_round_trip():
    rnd = random.Random(41)
    sk, pk = KeyGen(128, rnd)
    n = sk.N
    osk = R.SecretKey(N=n, Lambda=sk.Lambda)
    ms = [rnd.randrange(n) for _ in range(6)]
    for rand in (None, rnd):                                  # host CSPRNG and an injected source
        cts = pk.EncryptBatch(ms, rand)
        assert all(c.Level == ENC_LEVEL_ONE and c.EncMethod == REGULAR for c in cts)
        assert sk.DecryptBatch(cts) == ms == [R.decrypt(osk, R.Ciphertext(c.C)) for c in cts]
        m2 = [rnd.randrange(n * n) for _ in range(3)]
        assert sk.DecryptBatch(pk.EncryptAtLevelBatch(m2, ENC_LEVEL_TWO, rand)) == m2
        nested = pk.NestedEncryptBatch(ms, rand)             # TestNestedEncryptDecrypt, paillier_test.go:96-108
        assert all(c.Level == ENC_LEVEL_TWO for c in nested)
        assert sk.NestedDecryptBatch(nested) == ms == [R.nested_decrypt(osk, R.Ciphertext(c.C, ENC_LEVEL_TWO)) for c in nested]
    assert sk.DecryptBatch(pk.EncryptZeroBatch(3)) == [0, 0, 0]
    assert sk.DecryptBatch(pk.EncryptOneBatch(3)) == [1, 1, 1]
    assert sk.DecryptBatch(pk.EncryptZeroAtLevelBatch(2, ENC_LEVEL_TWO)) == [0, 0]
    assert sk.DecryptBatch(pk.EncryptOneAtLevelBatch(2, ENC_LEVEL_TWO)) == [1, 1]
    assert pk.EncryptBatch([]) == []
    # two draws of the same plaintext differ (fresh randomness per item)
    a, b = pk.EncryptBatch([5, 5])
    assert a.C != b.C
    sk.close(); pk.close()


def test_alt_encrypt_at_level_draws():
    rnd = random.Random(43)
    sk, pk = KeyGen(128, rnd)                                  # KeyGen sets H = r^2 mod n and K = 2^(secparam/2), paillier.go:151-166
    n = sk.N
    ms = [rnd.randrange(n) for _ in range(5)]
    for level, mm in ((ENC_LEVEL_ONE, ms), (ENC_LEVEL_TWO, [m * 3 + n for m in ms])):
        cts = pk.AltEncryptAtLevelBatch(mm, level)
        assert all(c.Level == level and c.EncMethod == ALTERNATIVE for c in cts)
        assert sk.DecryptBatch(cts) == mm
        assert [c.C for c in pk.AltEncryptAtLevelBatch(mm, level, rnd)] != [c.C for c in cts]
    sk.close(); pk.close()


def test_randomize_and_nested_randomize_draws():
    rnd = random.Random(47)
    sk, pk = KeyGen(128, rnd)
    n = sk.N
    opk = R.PublicKey(N=n)
    ms = [rnd.randrange(n) for _ in range(4)]
    cts = pk.EncryptBatch(ms, rnd)
    rz = pk.RandomizeBatch(cts)                                # TestRandomize: same plaintext, another ciphertext
    assert [c.C for c in rz] != [c.C for c in cts] and sk.DecryptBatch(rz) == ms
    assert all(c.EncMethod == MIXED for c in rz)
    # Randomize of a level-2 ciphertext multiplies by a LEVEL-1 Encrypt(0) modulo n^3 (operations.go:67-69 via Add)
    m2 = [rnd.randrange(n * n) for _ in range(3)]
    rs2 = pk._draw_units(3, rnd)
    c2 = pk.EncryptAtLevelBatch(m2, ENC_LEVEL_TWO, rnd)
    got = pk.RandomizeWithRBatch(c2, rs2)
    want = [R.add(opk, R.Ciphertext(c.C, ENC_LEVEL_TWO), R.encrypt_with_r(opk, 0, r)) for c, r in zip(c2, rs2)]
    assert [(g.C, g.Level) for g in got] == [(w.C, w.Level) for w in want]
    # NestedRandomize (operations_test.go:96-128): plaintext unchanged, (a, b) returned and reproducible
    nested = pk.NestedEncryptBatch(ms, rnd)
    out, As, Bs = pk.NestedRandomizeBatch(nested, rnd)
    assert sk.NestedDecryptBatch(out) == ms
    assert [c.C for c in out] == [R.nested_randomize_with(opk, R.Ciphertext(c.C, ENC_LEVEL_TWO), a, b).C for c, a, b in zip(nested, As, Bs)]
    sk.close(); pk.close()


def test_variadic_sub():
    rnd = random.Random(53)
    sk, pk = KeyGen(128, rnd)
    n = sk.N
    opk = R.PublicKey(N=n)
    ms = [rnd.randrange(n) for _ in range(7)]
    cts = pk.EncryptBatch(ms, rnd)
    for k in (1, 2, 3, 7):                                     # TestSub (operations_test.go:30-51) and longer argument lists
        got = pk.SubBatch(cts[:k])
        want = R.sub(opk, *[R.Ciphertext(c.C) for c in cts[:k]])
        assert (got.C, got.Level, got.EncMethod) == (want.C, want.Level, MIXED)
        assert sk.DecryptBatch([got]) == [(ms[0] - sum(ms[1:k])) % n]
    m2 = [rnd.randrange(n * n) for _ in range(3)]
    c2 = pk.EncryptAtLevelBatch(m2, ENC_LEVEL_TWO, rnd)
    got = pk.SubBatch(c2)
    assert got.Level == ENC_LEVEL_TWO and got.C == R.sub(opk, *[R.Ciphertext(c.C, ENC_LEVEL_TWO) for c in c2]).C
    assert sk.DecryptBatch([got]) == [(m2[0] - m2[1] - m2[2]) % (n * n)]
    with pytest.raises(PgpuError) as e:
        pk.SubBatch([cts[0], Ciphertext(n)])                   # n is not a unit mod n^2
    assert e.value.code == PGPU_ERR_NOT_INVERTIBLE
    sk.close(); pk.close()


def test_threshold_public_key_and_verify_partial_decryption():
    rnd = random.Random(59)
    keys = ThresholdKeyGenerator(64, 3, 2, rng=rnd, batch=512).GenerateKeys()
    for k in keys:
        k.VerifyPartialDecryption(count=2, rand=rnd)           # thresholdkey_test.go: TestVerifyPartialDecryption
    keys[0].VerifyPartialDecryption()                         # one check, host CSPRNG
    tpk = keys[0].PublicKey()
    assert not hasattr(tpk, "Share") and (tpk.N, tpk.Threshold, tpk.VerificationKeys) == (keys[0].N, 2, keys[0].VerificationKeys)
    n = tpk.N
    cts = tpk.EncryptBatch([11 % n, 12 % n], rnd)
    shares = [k.PartialDecryptionWithZKPBatch([c.C for c in cts], [rnd.randrange(n * n) for _ in cts]) for k in keys[:2]]
    assert tpk.CombinePartialDecryptionsZKPBatch(shares) == [11 % n, 12 % n]
    # a share-holder whose Share was corrupted fails its own check
    bad = type(keys[1])(keys[1].N, 3, 2, keys[1].VerificationKey, keys[1].VerificationKeys, keys[1].ID, keys[1].Share + 1)
    with pytest.raises(ValueError, match="Invalid share"):
        bad.VerifyPartialDecryption(rand=rnd)
    for k in keys + [tpk, bad]:
        k.close()


def test_cpp_mirror_callers():
    """the same callers through the C++ host mirror (tests/cpp/callers_test.cpp) on golden keys"""
    import json
    import os
    import subprocess
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_cpp_host_mirror", os.path.join(here, "test_cpp_host_mirror.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ROOT, build_callers_test = mod.ROOT, mod.build_callers_test
    V = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
    I = lambda s: int(s, 16)
    c = V["cases"]["paillier_64"]
    p, q = I(c["p"]), I(c["q"])
    lines = ["paillier", hex(p * q), hex((p - 1) * (q - 1)), c["H"], str(I(c["K"]).bit_length() - 1)]
    t = V["cases"]["threshold_512"]
    lines += ["threshold", hex(I(t["p"]) * I(t["q"])), str(t["l"]), str(t["w"]), t["V"]] + t["vi"]
    for i in range(t["l"]):
        lines += [str(i + 1), t["shares"][i]]
    # device-resident chaining (DeviceBuffer, *Dev methods) on the 64-bit key; a 1-of-1 threshold key for ThresholdGroup
    lines += ["device", hex(p * q), hex((p - 1) * (q - 1))]
    one = ThresholdKeyGenerator(512, 1, 1, rng=random.Random(5)).with_safe_primes(I(t["p"]), I(t["q"])).GenerateKeys()[0]
    lines += ["group", hex(one.N), hex(one.VerificationKey), hex(one.VerificationKeys[0]), hex(one.Share)]
    one.close()
    r = subprocess.run([build_callers_test()], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "callers ok" in r.stdout


def test_python_device_buffers_chain():
    """DeviceBuffer and the *Dev methods of the Python mirror (the same objects the Go binding and the C++ mirror have):
    Encrypt -> ConstMult -> AddPairs -> AddReduce -> Decrypt without leaving the GPU, against the host-buffer batch methods"""
    import numpy as np
    from paillier_b200 import synth
    from paillier_b200.api import SecretKey, from_records, to_records
    p, q = synth.load_key("paillier_2048")
    n = p * q
    sk = SecretKey(n, p=p, q=q)
    count = 40
    rnd = random.Random(61)
    ms = [rnd.randrange(n) for _ in range(count)]
    rs = sk._draw_units(count, rnd)
    ks = [rnd.getrandbits(64) | 1 for _ in range(count)]
    bm, br = sk.NewDeviceBuffer(count * sk.w_n).Upload(to_records(ms, sk.w_n)), sk.NewDeviceBuffer(count * sk.w_n).Upload(to_records(rs, sk.w_n))
    bk = sk.NewDeviceBuffer(count * 8).Upload(np.array(ks, dtype=np.uint64))
    bc, bc2, bt, bo = sk.NewDeviceBuffer(count * sk.w_n2), sk.NewDeviceBuffer(count * sk.w_n2), sk.NewDeviceBuffer(sk.w_n2), sk.NewDeviceBuffer(sk.w_n)
    assert len(bc) == count * sk.w_n2
    sk.EncryptWithRDev(count, bm, br, bc)
    sk.ConstMultDev(count, bc, bk, 8, bc2)          # k_i * m_i
    sk.AddPairsDev(count, bc2, bc, bc2)             # + m_i
    sk.AddReduceDev(count, bc2, bt)
    sk.DecryptDev(1, bt, bo)
    sk.Sync()
    cts = sk.EncryptWithRBatch(ms, rs)
    assert from_records(bc.Download(), sk.w_n2) == [c.C for c in cts]
    assert from_records(bo.Download(), sk.w_n) == [sum((k + 1) * m for k, m in zip(ks, ms)) % n]
    assert from_records(bc.Download(sk.w_n2, offset=3 * sk.w_n2), sk.w_n2) == [cts[3].C]
    with pytest.raises(ValueError):
        sk.EncryptWithRDev(count + 1, bm, br, bc)   # a batch larger than its buffers
    for b in (bm, br, bk, bc, bc2, bt, bo):
        b.Free()
    sk.close()
