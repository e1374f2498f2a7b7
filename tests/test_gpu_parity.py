"""GPU parity: libpaillier_b200.so (through the C ABI) against the oracle on identical inputs.
Bit-exact: all outputs are canonical residues."""
import numpy as np
import pytest

from oracle import gmp_ref as G
from oracle import paillier_ref as R
from paillier_b200 import synth
from paillier_b200.api import (Ciphertext, PublicKey, SecretKey, ThresholdSecretKey, from_records, to_records,
                               MOD_N, MOD_N2, MOD_N3)

pytestmark = pytest.mark.gpu


def _key(name):
    p, q = synth.load_key(name)
    return p, q, p * q


@pytest.fixture(scope="module")
def sk2048():
    p, q, n = _key("paillier_2048")
    sk = SecretKey(n, p=p, q=q)
    yield sk, R.keygen_from_primes(p, q)[0]
    sk.close()


# ---- reference KATs through the GPU path -------------------------------------------------

def test_partial_decrypt_kat_on_gpu():
    # thresholdkey_test.go:58-74 (TestDecrypt): 56^(862*2*10!) mod (101*103)^2 = 40644522
    tsk = ThresholdSecretKey(101 * 103, 10, 6, VerificationKey=4, VerificationKeys=[], ID=9, Share=862)
    pd = tsk.PartialDecryptBatch([56])
    assert pd[0].ID == 9 and pd[0].Decryption == 40644522
    tsk.close()


def test_exp_kats_on_gpu():
    # thresholdkey_generator_test.go:314-324: 54^(s*10!) mod 101^2 for s = 12, 90, 103
    pk = PublicKey(101)
    d = R.factorial(10)
    assert pk.ExpBatch([54, 54, 54], [12 * d, 90 * d, 103 * d]) == [6162, 304, 2728]
    # thresholdkey_test.go:32-46 (TestExp), modulus 49 = 7^2
    pk7 = PublicKey(7)
    assert pk7.ExpBatch([720 % 49, 720 % 49], [10, 0]) == [43, 1]
    assert pk7.ExpSharedBatch([720 % 49], 10) == [43]
    assert pk7.ExpSharedBatch([720 % 49], 0) == [1]
    pk.close(); pk7.close()


# ---- encrypt / decrypt ----------------------------------------------------------------------

@pytest.mark.parametrize("name", ["paillier_64", "paillier_1024", "paillier_2048"])
def test_encrypt_decrypt_parity_python_oracle(name):
    p, q, n = _key(name)
    osk, opk = R.keygen_from_primes(p, q)
    sk = SecretKey(n, Lambda=osk.Lambda)        # p, q recovered from lambda as the Go SecretKey only has Lambda
    ms = from_records(synth.plaintexts(12, n, sk.w_n), sk.w_n)
    rs = from_records(synth.randomness(12, n, sk.w_n), sk.w_n)
    ms[0], ms[1], ms[2], rs[3], rs[4] = 0, n - 1, 1, 1, n - 1
    cts = sk.EncryptWithRBatch(ms, rs)                               # secret-key path (CRT over p^2, q^2)
    assert [c.C for c in cts] == [R.encrypt_with_r(opk, m, r).C for m, r in zip(ms, rs)]
    pub = from_records(PublicKey.encrypt_with_r_records(sk, to_records(ms, sk.w_n), to_records(rs, sk.w_n)), sk.w_n2)
    assert pub == [c.C for c in cts]                                 # public-key path (r^n mod n^2)
    assert sk.DecryptBatch(cts) == ms == [R.decrypt(osk, R.Ciphertext(c.C)) for c in cts]
    sk.close()


@pytest.mark.parametrize("name", ["paillier_64", "paillier_1024", "paillier_2048", "threshold_512", "threshold_3072"])
def test_secret_key_encrypt_edge_randomness(name):
    """EncryptWithR over p^2, q^2 with r values that stress the recombination: multiples of p and q (r^n = 0 mod p^2),
    r on either side of p^2 and q^2, r = 0 (the reference does not reject it: c = 0), against both oracles and the
    public-key path."""
    p, q, n = _key(name)
    osk, opk = R.keygen_from_primes(p, q)
    sk = SecretKey(n, p=p, q=q)
    lo2, hi2 = min(p * p, q * q), max(p * p, q * q)
    rs = [0, 1, 2, p, q, 2 * p, 3 * q, p * (q - 1), q * (p - 1), n - 1, n - 2,
          lo2 % n, (lo2 - 1) % n, (lo2 + 1) % n, hi2 % n, (hi2 - 1) % n, (hi2 + 1) % n, n // 2, n // 3]
    ms = [(i * 0x9E3779B97F4A7C15 + 7) % n for i in range(len(rs))]
    ms[0], ms[1], ms[2] = 0, n - 1, 1
    mrec, rrec = to_records(ms, sk.w_n), to_records(rs, sk.w_n)
    got = from_records(sk.encrypt_with_r_records(mrec, rrec), sk.w_n2)
    assert got == [R.encrypt_with_r(opk, m, r).C for m, r in zip(ms, rs)]
    assert got == from_records(PublicKey.encrypt_with_r_records(sk, mrec, rrec), sk.w_n2)
    assert got == from_records(G.encrypt_with_r(n, mrec, rrec, sk.w_n), sk.w_n2)
    sk.close()


@pytest.mark.parametrize("name", ["paillier_64", "paillier_2048"])
def test_offline_online_encrypt_matches_encrypt_with_r(name):
    """r^n pool: EncryptWithRn(m, r^n mod n^2) == EncryptWithR(m, r) (paillier.go:206-218), through the
    public-key and the secret-key precomputation, against the Python oracle."""
    p, q, n = _key(name)
    osk, opk = R.keygen_from_primes(p, q)
    sk = SecretKey(n, p=p, q=q)
    pk = PublicKey(n)
    ms = from_records(synth.plaintexts(40, n, sk.w_n), sk.w_n)
    rs = from_records(synth.randomness(40, n, sk.w_n), sk.w_n)
    ms[0], ms[1], rs[2], rs[3] = 0, n - 1, 1, n - 1
    pool_pk, pool_sk = pk.PrecomputeRnBatch(rs), sk.PrecomputeRnBatch(rs)
    assert pool_pk == pool_sk == [pow(r, n, n * n) for r in rs]
    want = [R.encrypt_with_r(opk, m, r).C for m, r in zip(ms, rs)]
    assert [c.C for c in pk.EncryptWithRnBatch(ms, pool_pk)] == want
    assert [c.C for c in sk.EncryptWithRnBatch(ms, pool_sk)] == want
    assert pk.EncryptWithRnBatch([], []) == []
    sk.close(); pk.close()


def test_empty_and_single_batches(sk2048):
    sk, _ = sk2048
    assert sk.EncryptWithRBatch([], []) == []
    assert sk.DecryptBatch([]) == []
    c = sk.EncryptWithRBatch([42], [3])
    assert sk.DecryptBatch(c) == [42]


def test_encrypt_decrypt_2048_large_batch_vs_libgmp(sk2048):
    # ragged count: more items than resident groups and not a multiple of anything
    sk, osk = sk2048
    n, count = sk.N, 20011
    m = synth.plaintexts(count, n, sk.w_n)
    r = synth.randomness(count, n, sk.w_n)
    c = PublicKey.encrypt_with_r_records(sk, m, r)
    assert np.array_equal(sk.encrypt_with_r_records(m, r), c)        # secret-key path over the same ragged batch
    nref = 4096
    ref = G.encrypt_with_r(n, m[:nref * sk.w_n], r[:nref * sk.w_n], sk.w_n)
    assert np.array_equal(c[:nref * sk.w_n2], ref)
    # tail of the batch against libgmp as well (last partially filled round of groups)
    tail = slice((count - 64) * sk.w_n, count * sk.w_n)
    ref_tail = G.encrypt_with_r(n, m[tail], r[tail], sk.w_n)
    assert np.array_equal(c[(count - 64) * sk.w_n2:], ref_tail)
    # CRT decrypt: round trip over the whole batch + libgmp lambda-decrypt on a subset
    d = sk.decrypt_records(c)
    assert np.array_equal(d, m)
    refd = G.decrypt(n, osk.Lambda, c[:512 * sk.w_n2], sk.w_n)
    assert np.array_equal(d[:512 * sk.w_n], refd)


# ---- homomorphic operations -------------------------------------------------------------------

def test_const_mult_add_dot(sk2048):
    sk, osk = sk2048
    n, n2 = sk.N, sk.N ** 2
    count = 300
    m = synth.plaintexts(count, n, sk.w_n)
    r = synth.randomness(count, n, sk.w_n)
    c = sk.encrypt_with_r_records(m, r)
    k = synth.scalars_u64(count)
    k[0], k[1], k[2] = 0, 1, 2 ** 64 - 1
    out = sk.const_mult_records(c, k.view(np.uint8), 8)
    ref = G.modexp(n2, c, sk.w_n2, k.view(np.uint8), 8)
    assert np.array_equal(out, ref)
    assert from_records(out[:sk.w_n2], sk.w_n2) == [1]          # ConstMult by 0 -> 1 (gmp Exp semantics)
    # Add over the batch == left fold of the reference
    s = sk.add_reduce_records(c)
    assert np.array_equal(s, G.add_reduce(n2, c, sk.w_n2))
    ms = from_records(m, sk.w_n)
    assert sk.DecryptBatch([Ciphertext(from_records(s, sk.w_n2)[0])]) == [sum(ms) % n]
    # pairs
    pr = sk.add_pairs_records(c, out)
    assert np.array_equal(pr, G.modmul(n2, c, out, sk.w_n2))
    # encrypted dot product (BASELINE config 3) == Add(ConstMult(c_i, k_i))
    dot = sk.dot_u64_records(c, k)
    assert np.array_equal(dot, G.add_reduce(n2, ref, sk.w_n2))
    expect = sum(int(ki) * mi for ki, mi in zip(k, ms)) % n
    assert sk.DecryptBatch([Ciphertext(from_records(dot, sk.w_n2)[0])]) == [expect]


@pytest.mark.parametrize("count", [0, 1, 2, 17, 1000, 9473, 40000])
def test_add_reduce_sizes(sk2048, count):
    sk, _ = sk2048
    n2 = sk.N ** 2
    c = synth.random_records(count, sk.w_n2, 4094, stream=9)
    got = sk.add_reduce_records(c)
    assert np.array_equal(got, G.add_reduce(n2, c, sk.w_n2)) if count else from_records(got, sk.w_n2) == [1]


def test_generic_modexp_n2_n3(sk2048):
    sk, _ = sk2048
    n2, n3 = sk.N ** 2, sk.N ** 3
    count = 24
    for modsel, mod, width, bits in ((MOD_N2, n2, sk.w_n2, 4094), (MOD_N3, n3, sk.w_n3, 6140)):
        base = synth.random_records(count, width, bits, stream=11)
        for eb in (4, 8, 256, 512):
            exp = synth.random_records(count, eb, eb * 8, stream=12)
            got = sk.modexp_records(modsel, base, exp, eb, width)
            assert np.array_equal(got, G.modexp(mod, base, width, exp, eb)), (modsel, eb)
        bases = from_records(base, width)
        e = sk.N
        assert sk.ExpSharedBatch(bases[:4], e, modsel) == [pow(b, e, mod) for b in bases[:4]]
        got = sk.modmul_records(modsel, base, base[::-1].copy(), width)
        assert np.array_equal(got, G.modmul(mod, base, base[::-1].copy(), width))


# ---- threshold partial decryption ----------------------------------------------------------------

@pytest.mark.parametrize("name,l,w", [("threshold_512", 5, 3), ("threshold_2048", 8, 5), ("threshold_3072", 8, 5)])
def test_partial_decrypt_parity(name, l, w):
    p, q, n = _key(name)
    import random
    rnd = random.Random(synth.SEED)
    nm = n * ((p - 1) // 2) * ((q - 1) // 2)
    keys = R.threshold_keys_from(p, q, l, w, v_seed=rnd.randrange(2, n * n), coeffs=[rnd.randrange(nm) for _ in range(w - 1)])
    ok = keys[2]
    tsk = ThresholdSecretKey(n, l, w, ok.VerificationKey, ok.VerificationKeys, ok.ID, ok.Share)
    count = 200
    m = synth.plaintexts(count, n, tsk.w_n)
    r = synth.randomness(count, n, tsk.w_n)
    c = tsk.encrypt_with_r_records(m, r)
    out = tsk.partial_decrypt_records(c)
    ref = G.partial_decrypt(n, ok.Share, l, c, tsk.w_n2)
    assert np.array_equal(out, ref)
    c0 = from_records(c[:tsk.w_n2], tsk.w_n2)[0]
    assert from_records(out[:tsk.w_n2], tsk.w_n2)[0] == R.partial_decrypt(ok, c0).Decryption
    tsk.close()


def test_threshold_keygen_matches_oracle():
    # thresholdkey_generator.go:47-55 with injected randomness; verification keys computed on the GPU
    import random
    from paillier_b200.keygen import ThresholdKeyGenerator
    p, q, n = _key("threshold_512")
    keys = ThresholdKeyGenerator(512, 6, 4, rng=random.Random(99)).with_safe_primes(p, q).GenerateKeys()
    rnd = random.Random(99)
    n2, nm = n * n, n * ((p - 1) // 2) * ((q - 1) // 2)
    while True:
        r = rnd.randrange(1, n2)
        if r % p and r % q:
            break
    coeffs = [rnd.randrange(nm) for _ in range(3)]
    okeys = R.threshold_keys_from(p, q, 6, 4, v_seed=r, coeffs=coeffs)
    for k, ok in zip(keys, okeys):
        assert (k.ID, k.Share, k.VerificationKey, k.VerificationKeys) == (ok.ID, ok.Share, ok.VerificationKey, ok.VerificationKeys)
        k.close()


# ---- Sub / ModInverse ---------------------------------------------------------------------------------

def test_sub_and_modinverse(sk2048):
    sk, osk = sk2048
    n, n2 = sk.N, sk.N ** 2
    opk = R.PublicKey(N=n)
    ms = [5, 1000, n - 3, 77]
    rs = from_records(synth.randomness(8, n, sk.w_n), sk.w_n)
    a = sk.EncryptWithRBatch(ms, rs[:4])
    b = sk.EncryptWithRBatch([2, 1, 4, 70], rs[4:])
    got = sk.SubPairs(a, b)
    assert [g.C for g in got] == [R.sub(opk, R.Ciphertext(x.C), R.Ciphertext(y.C)).C for x, y in zip(a, b)]
    assert sk.DecryptBatch(got) == [3, 999, n - 7, 7]
    xs = [x.C for x in a] + [1, n2 - 1, 2]
    assert sk.ModInverseBatch(xs) == [pow(x, -1, n2) for x in xs]
    from paillier_b200._lib import PgpuError, PGPU_ERR_NOT_INVERTIBLE
    p, _, _ = _key("paillier_2048")
    with pytest.raises(PgpuError) as ei:
        sk.ModInverseBatch([3, p * 5, 7])
    assert ei.value.code == PGPU_ERR_NOT_INVERTIBLE and "item 1" in str(ei.value)
    n3 = sk.N ** 3
    ys = [pow(7, i + 1, n3) for i in range(3)]
    assert sk.ModInverseBatch(ys, MOD_N3) == [pow(y, -1, n3) for y in ys]


# ---- threshold ZKP and share combining (thresholdkey.go:149-326) --------------------------------------

def _threshold_setup(name, l, w, bits):
    import random
    from paillier_b200.keygen import ThresholdKeyGenerator
    p, q, n = _key(name)
    keys = ThresholdKeyGenerator(bits, l, w, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys()
    okeys = [R.ThresholdSecretKey(N=n, TotalNumberOfDecryptionServers=l, Threshold=w, VerificationKey=k.VerificationKey,
                                  VerificationKeys=k.VerificationKeys, ID=k.ID, Share=k.Share) for k in keys]
    return n, keys, okeys


@pytest.mark.parametrize("name,l,w,bits,count", [("threshold_512", 6, 4, 512, 40), ("threshold_2048", 8, 5, 2048, 6)])
def test_zkp_prove_verify_combine(name, l, w, bits, count):
    n, keys, okeys = _threshold_setup(name, l, w, bits)
    n2 = n * n
    tk = keys[0]
    ms = from_records(synth.plaintexts(count, n, tk.w_n), tk.w_n)
    ms[0] = 0
    cs = [c.C for c in tk.EncryptWithRBatch(ms, from_records(synth.randomness(count, n, tk.w_n), tk.w_n))]
    rs = from_records(synth.random_records(count, tk.w_n2, n2.bit_length() - 1, stream=21), tk.w_n2)
    rs[1] = 0
    proofs = []
    for k, ok in zip(keys[:w + 1], okeys):
        zk = k.PartialDecryptionWithZKPBatch(cs, rs)
        for i in (0, 1, count - 1):      # transcript bit-exact with the oracle for fixed randomness
            o = R.partial_decryption_with_zkp(ok, cs[i], rs[i])
            assert (zk[i].ID, zk[i].Decryption, zk[i].E, zk[i].Z) == (o.ID, o.Decryption, o.E, o.Z)
        proofs.append(zk)
    # verification: accept honest proofs, reject a tampered one and a wrong server id (thresholdkey_test.go:137-149,283-292)
    assert all(tk.VerifyProofBatch(proofs[2]))
    bad = list(proofs[3])
    bad[2] = type(bad[2])(bad[2].ID, bad[2].Decryption, bad[2].E, bad[2].Z + 1, bad[2].C)
    res = tk.VerifyProofBatch(bad)
    assert res[2] is False and all(r for i, r in enumerate(res) if i != 2)
    wrong_id = [type(p)(p.ID + 1, p.Decryption, p.E, p.Z, p.C) for p in proofs[0]]
    assert not any(tk.VerifyProofBatch(wrong_id))
    from paillier_b200.api import PartialDecryption
    shares = [[PartialDecryption(p.ID, p.Decryption) for p in s] for s in proofs]
    assert tk.CombinePartialDecryptionsBatch(shares[:w]) == ms
    assert tk.CombinePartialDecryptionsBatch(shares) == ms
    assert tk.CombinePartialDecryptionsBatch(shares[1:w + 1][::-1]) == ms
    otk = R.threshold_public_key(okeys[0])
    assert tk.CombinePartialDecryptionsBatch(shares[:w])[0] == R.combine_partial_decryptions(
        otk, [R.PartialDecryption(s[0].ID, s[0].Decryption) for s in shares[:w]])
    assert tk.CombinePartialDecryptionsZKPBatch(proofs) == ms
    from paillier_b200._lib import PgpuError, PGPU_ERR_THRESHOLD
    with pytest.raises(PgpuError) as ei:
        tk.CombinePartialDecryptionsBatch(shares[:w - 1])          # "Threshold not meet"
    assert ei.value.code == PGPU_ERR_THRESHOLD
    with pytest.raises(PgpuError) as ei:
        tk.CombinePartialDecryptionsBatch(shares[:w - 1] + [shares[0]])   # duplicate server
    assert ei.value.code == PGPU_ERR_THRESHOLD
    for k in keys:
        k.close()


def test_combine_kat_on_gpu():
    # thresholdkey_test.go:267-281 (TestDecryption): two literal shares -> 100
    from paillier_b200.api import ThresholdPublicKey, PartialDecryption
    tk = ThresholdPublicKey(637753, 2, 2, 70661107826, [])
    got = tk.CombinePartialDecryptionsBatch([[PartialDecryption(1, 384111638639)], [PartialDecryption(2, 235243761043)]])
    assert got == [100]
    tk.close()


def test_verify_kats_on_gpu():
    # thresholdkey_test.go:109-135 (TestVerifyPart1 / TestVerifyPart2): n = 131
    pk = PublicKey(131)
    n2 = 131 * 131
    c4, d2 = 99 ** 4 % n2, 101 ** 2 % n2
    a1 = pk.ExpBatch([c4], [88])[0]
    a2 = pk.ModInverseBatch(pk.ExpBatch([d2], [112]))[0]
    assert pk.MulModBatch([a1], [a2]) == [11986]
    b1 = pk.ExpBatch([101], [88])[0]
    b2 = pk.ModInverseBatch(pk.ExpBatch([77], [112]))[0]
    assert pk.MulModBatch([b1], [b2]) == [14602]
    pk.close()


def test_batch_inversion_path(sk2048):
    # >= 128 items take Montgomery's batch-inversion route (chunks of 32 + a ragged tail)
    sk, _ = sk2048
    n2, n3 = sk.N ** 2, sk.N ** 3
    import random
    rnd = random.Random(77)
    for modsel, mod, count in ((MOD_N2, n2, 300), (MOD_N3, n3, 131)):
        xs = [rnd.randrange(1, mod) for _ in range(count)]
        xs[0], xs[1], xs[-1] = 1, mod - 1, 2
        assert sk.ModInverseBatch(xs, modsel) == [pow(x, -1, mod) for x in xs]
    p, _, _ = _key("paillier_2048")
    from paillier_b200._lib import PgpuError, PGPU_ERR_NOT_INVERTIBLE
    xs = [rnd.randrange(1, n2) for _ in range(200)]
    xs[137] = p * 12345
    xs[190] = 0
    with pytest.raises(PgpuError) as ei:
        sk.ModInverseBatch(xs)
    assert ei.value.code == PGPU_ERR_NOT_INVERTIBLE and "item 137" in str(ei.value)


def test_zkp_large_batch_fixed_base_and_batch_inverse():
    # 160 proofs: V^r / V^Z through the comb table, the verifier's inversions through the batch route
    n, keys, okeys = _threshold_setup("threshold_512", 5, 3, 512)
    tk = keys[1]
    count = 160
    ms = from_records(synth.plaintexts(count, n, tk.w_n), tk.w_n)
    cs = [c.C for c in tk.EncryptWithRBatch(ms, from_records(synth.randomness(count, n, tk.w_n), tk.w_n))]
    rs = from_records(synth.random_records(count, tk.w_n2, (n * n).bit_length() - 1, stream=23), tk.w_n2)
    zk = tk.PartialDecryptionWithZKPBatch(cs, rs)
    for i in (0, 77, count - 1):
        o = R.partial_decryption_with_zkp(okeys[1], cs[i], rs[i])
        assert (zk[i].Decryption, zk[i].E, zk[i].Z) == (o.Decryption, o.E, o.Z)
    bad = list(zk)
    bad[99] = type(bad[99])(bad[99].ID, bad[99].Decryption, bad[99].E ^ 1, bad[99].Z, bad[99].C)
    res = tk.VerifyProofBatch(bad)
    assert res[99] is False and sum(res) == count - 1
    for k in keys:
        k.close()


def test_dot_product_pippenger_large_batch(sk2048):
    # >= 16 items per resident group take the bucket method; it must equal ConstMult + Add bit for bit
    import os
    sk, _ = sk2048
    n = sk.N
    count = 170_001
    m = synth.plaintexts(count, n, sk.w_n)
    c = sk.encrypt_with_r_records(m, synth.randomness(count, n, sk.w_n))
    k = synth.scalars_u64(count)
    k[0], k[1], k[2], k[-1] = 0, 1, 2 ** 64 - 1, 0
    fast = sk.dot_u64_records(c, k)
    os.environ["PGPU_DOT_SIMPLE"] = "1"
    try:
        simple = sk.dot_u64_records(c, k)
    finally:
        del os.environ["PGPU_DOT_SIMPLE"]
    assert np.array_equal(fast, simple)
    ms = from_records(m, sk.w_n)
    expect = sum(int(ki) * mi for ki, mi in zip(k, ms)) % n
    assert sk.DecryptBatch([Ciphertext(from_records(fast, sk.w_n2)[0])]) == [expect]
    # a subset against the libgmp oracle
    sub = 2048
    ref = G.add_reduce(n * n, G.modexp(n * n, c[:sub * sk.w_n2], sk.w_n2, k[:sub].view(np.uint8), 8), sk.w_n2)
    assert np.array_equal(sk.dot_u64_records(c[:sub * sk.w_n2], k[:sub]), ref)


def test_encrypt_batch_draws_units(sk2048):
    # PublicKey.Encrypt (paillier.go:192,258-269): r drawn on the host, unit check batched on the GPU
    import random
    sk, _ = sk2048
    ms = [0, 1, sk.N - 1, 424242]
    cts = sk.EncryptBatch(ms, rand=random.Random(3))
    assert sk.DecryptBatch(cts) == ms
    assert len({c.C for c in cts}) == 4
    assert sk.ModInverseBatch([3, sk.N - 1], MOD_N) == [pow(3, -1, sk.N), sk.N - 1]


def test_carry_chain_edge_values():
    # operands made of all-ones / all-zero limbs and values next to the modulus, through every multiplier shape
    # (32, 64, 96, 128, 192 limbs): squarings (dedicated kernel for the 32- and 64-limb shapes) and multiplications
    import random
    rnd = random.Random(99)
    for name in ("paillier_64", "paillier_1024", "paillier_2048", "threshold_3072"):
        p, q, n = _key(name)
        pk = PublicKey(n)
        for modsel, mod, width in ((MOD_N2, n * n, pk.w_n2), (MOD_N3, n ** 3, pk.w_n3)):
            if not width:
                continue
            limbs = width // 4
            vals = [0, 1, 2, mod - 1, mod - 2, mod >> 1, (mod >> 1) + 1, (1 << (32 * (limbs - 1))) % mod]
            vals += [((1 << b) - 1) % mod for b in (31, 32, 33, 64, 32 * limbs // 2, 32 * limbs // 2 + 1, mod.bit_length() - 1)]
            vals += [int("ffffffff00000000" * (limbs // 2), 16) % mod, int("00000000ffffffff" * (limbs // 2), 16) % mod]
            vals += [(mod - (1 << b)) % mod for b in (0, 32, 64, 32 * limbs // 4)]
            vals += [rnd.randrange(mod) for _ in range(6)]
            assert pk.ExpSharedBatch(vals, 2, modsel) == [v * v % mod for v in vals]
            assert pk.ExpSharedBatch(vals, 65537, modsel) == [pow(v, 65537, mod) for v in vals]
            rev = vals[::-1]
            assert pk.MulModBatch(vals, rev, modsel) == [a * b % mod for a, b in zip(vals, rev)]
        pk.close()
    # the CRT halves of Decrypt run the 64-limb squaring kernel: plaintexts and ciphertexts at the extremes
    p, q, n = _key("paillier_2048")
    sk = SecretKey(n, p=p, q=q)
    ms = [0, 1, n - 1, (1 << 2047) % n, (1 << 1024) - 1, n >> 1]
    for r in (1, n - 1, 2):
        assert sk.DecryptBatch(sk.EncryptWithRBatch(ms, [r] * len(ms))) == ms
    sk.close()


def test_combine_zkp_filters_per_ciphertext():
    # CombinePartialDecryptionsZKP (thresholdkey.go:164-172) drops a share PER CIPHERTEXT: one bad proof must not remove its
    # server from the other ciphertexts.  6 servers, threshold 4, 12 ciphertexts:
    #   ciphertext 2: server 1's proof tampered          -> combined from the 5 others
    #   ciphertext 5: servers 2 and 3 tampered           -> combined from 4
    #   ciphertext 7: servers 1, 2, 3 tampered           -> 3 < threshold: "Threshold not meet" for this one only
    from paillier_b200._lib import PgpuError, PGPU_ERR_THRESHOLD
    n, keys, okeys = _threshold_setup("threshold_512", 6, 4, 512)
    tk = keys[0]
    count = 12
    ms = from_records(synth.plaintexts(count, n, tk.w_n), tk.w_n)
    cs = [c.C for c in tk.EncryptWithRBatch(ms, from_records(synth.randomness(count, n, tk.w_n), tk.w_n))]
    rs = from_records(synth.random_records(count, tk.w_n2, (n * n).bit_length() - 1, stream=23), tk.w_n2)
    proofs = [list(k.PartialDecryptionWithZKPBatch(cs, rs)) for k in keys]
    def tamper(server, item):
        p = proofs[server][item]
        proofs[server][item] = type(p)(p.ID, p.Decryption, p.E, p.Z + 1, p.C)
    tamper(0, 2); tamper(1, 5); tamper(2, 5); tamper(0, 7); tamper(1, 7); tamper(2, 7)
    got = tk.CombinePartialDecryptionsZKPBatch(proofs, strict=False)
    assert got[7] is None
    assert [g for i, g in enumerate(got) if i != 7] == [m for i, m in enumerate(ms) if i != 7]
    with pytest.raises(PgpuError) as ei:
        tk.CombinePartialDecryptionsZKPBatch(proofs)
    assert ei.value.code == PGPU_ERR_THRESHOLD and "1 of 12" in str(ei.value)
    # the oracle, ciphertext by ciphertext
    otk = R.threshold_public_key(okeys[0])
    for i in (2, 5):
        shares = [R.PartialDecryptionZKP(ID=p[i].ID, Decryption=p[i].Decryption, Key=otk, E=p[i].E, Z=p[i].Z, C=p[i].C) for p in proofs]
        assert R.combine_partial_decryptions_zkp(otk, shares) == ms[i]
    for k in keys:
        k.close()


def test_dedicated_squaring_is_race_free_by_determinism(sk2048):
    # compute-sanitizer is closed on the GPU pool (profiles/r02_sanitizer_unavailable.txt), so the shared-memory exchange of
    # Mont::sqr (block products handed between the lanes of a group, __syncwarp() fences; 64-limb moduli = CRT Decrypt) is
    # checked the other way: the same batch repeatedly, batches that leave 1, 2, ... 8 groups of a warp active, and the exchange
    # switched off in a child process (PGPU_NO_SQR=1) must all give the bits libgmp gives.
    import os
    import subprocess
    import sys
    sk, osk = sk2048
    n = sk.N
    lam = osk.Lambda
    count = 3001
    m = synth.plaintexts(count, n, sk.w_n)
    c = sk.encrypt_with_r_records(m, synth.randomness(count, n, sk.w_n))
    ref = G.decrypt(n, lam, c[:512 * sk.w_n2], sk.w_n)
    first = sk.decrypt_records(c)
    assert np.array_equal(first, m) and np.array_equal(first[:512 * sk.w_n], ref)
    for _ in range(6):
        assert np.array_equal(sk.decrypt_records(c), first)
    for k in (1, 2, 3, 5, 7, 8, 9, 31, 33):
        assert np.array_equal(sk.decrypt_records(c[:k * sk.w_n2]), m[:k * sk.w_n])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    child = ("import sys; sys.path.insert(0, %r)\n"
             "import numpy as np\nfrom paillier_b200 import synth\nfrom paillier_b200.api import SecretKey\n"
             "p, q = synth.load_key('paillier_2048'); sk = SecretKey(p * q, p=p, q=q)\n"
             "m = synth.plaintexts(3001, sk.N, sk.w_n); c = sk.encrypt_with_r_records(m, synth.randomness(3001, sk.N, sk.w_n))\n"
             "assert np.array_equal(sk.decrypt_records(c), m)\nprint('ok')\n" % root)
    r = subprocess.run([sys.executable, "-c", child], env=dict(os.environ, PGPU_NO_SQR="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_fused_prove_equals_separate_exponentiations():
    # PartialDecryptionWithZKP computes c^(2*delta*s) and (c^4)^r in one launch with shared squarings; with
    # PGPU_NO_FUSED_PROVE=1 (child process) it runs the two exponentiations separately: same (c_i, E, Z), and both equal libgmp's
    import hashlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = ("import sys, hashlib, random; sys.path.insert(0, %r)\n"
            "from paillier_b200 import synth\nfrom paillier_b200.keygen import ThresholdKeyGenerator\n"
            "p, q = synth.load_key('threshold_2048'); n = p * q\n"
            "tsk = ThresholdKeyGenerator(2048, 8, 5, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys()[3]\n"
            "c = tsk.encrypt_with_r_records(synth.plaintexts(700, n, tsk.w_n), synth.randomness(700, n, tsk.w_n))\n"
            "r = synth.random_records(700, tsk.w_n2, 2 * n.bit_length() - 2, stream=44)\n"
            "print('DIGEST', hashlib.sha256(b''.join(x.tobytes() for x in tsk.zkp_prove_records(c, r))).hexdigest())\n" % root)
    digests = []
    for env in ({}, {"PGPU_NO_FUSED_PROVE": "1"}):
        r = subprocess.run([sys.executable, "-c", prog], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        digests.append([l for l in r.stdout.splitlines() if l.startswith("DIGEST")][-1])
    assert digests[0] == digests[1]
    import random
    p, q, n = _key("threshold_2048")
    from paillier_b200.keygen import ThresholdKeyGenerator
    keys = ThresholdKeyGenerator(2048, 8, 5, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys()
    tsk = keys[3]
    for k in keys:
        if k is not tsk:
            k.close()
    c = tsk.encrypt_with_r_records(synth.plaintexts(700, n, tsk.w_n), synth.randomness(700, n, tsk.w_n))
    r = synth.random_records(700, tsk.w_n2, 2 * n.bit_length() - 2, stream=44)
    ref = G.pdec_zkp(n, tsk.Share, 8, tsk.VerificationKey, c, r, tsk.w_n2, tsk.w_z)
    assert "DIGEST " + hashlib.sha256(b"".join(x.tobytes() for x in ref)).hexdigest() == digests[0]
    tsk.close()
