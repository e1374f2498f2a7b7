"""The oracle against every known-answer test the reference holds for the hot path
(SURVEY.md section 4 / 8c).  Each case cites the reference test it restates."""
import pytest

from oracle import paillier_ref as R


def test_exp_semantics():
    # thresholdkey_test.go:32-46 (TestExp)
    assert R.tk_exp(720, 10, 49) == 43
    assert R.tk_exp(720, 0, 49) == 1
    assert R.tk_exp(720, -10, 49) == 8
    # ncw/gmp: exponent <= 0 -> 1 regardless of modulus; nil modulus -> plain power
    assert R.gmp_exp(5, -3, 7) == 1
    assert R.gmp_exp(3, 4, None) == 81


def test_delta_and_factorial():
    # thresholdkey_test.go:24-30 (TestDelta), utils_test.go:58-62 (TestFactorial)
    assert R.ThresholdPublicKey(N=1, TotalNumberOfDecryptionServers=6).delta() == 720
    assert R.factorial(6) == 720


def test_combine_shares_constant():
    # thresholdkey_test.go:48-56
    tk = R.ThresholdPublicKey(N=101 * 103, TotalNumberOfDecryptionServers=6)
    assert tk.combine_shares_constant() == 4558


def test_partial_decrypt_kat():
    # thresholdkey_test.go:58-74 (TestDecrypt)
    key = R.ThresholdSecretKey(N=101 * 103, TotalNumberOfDecryptionServers=10, Share=862, ID=9)
    pd = R.partial_decrypt(key, 56)
    assert pd.ID == 9 and pd.Decryption == 40644522


def test_verify_part1():
    # thresholdkey_test.go:109-121
    pd = R.PartialDecryptionZKP(ID=0, Decryption=101, Key=R.ThresholdPublicKey(N=131), E=112, Z=88, C=99)
    assert R.verify_part1(pd) == 11986


def test_verify_part2():
    # thresholdkey_test.go:123-135
    key = R.ThresholdPublicKey(N=131, VerificationKey=101, VerificationKeys=[77, 67])
    pd = R.PartialDecryptionZKP(ID=1, Decryption=0, Key=key, E=112, Z=88, C=0)
    assert R.verify_part2(pd) == 14602


def test_verify_partial_decryptions_rules():
    # thresholdkey_test.go:151-166
    tk = R.ThresholdPublicKey(N=1, Threshold=2)
    with pytest.raises(R.ThresholdError):
        R.verify_partial_decryptions(tk, [])
    R.verify_partial_decryptions(tk, [R.PartialDecryption(0, 0), R.PartialDecryption(1, 0)])
    with pytest.raises(R.ThresholdError):
        R.verify_partial_decryptions(tk, [R.PartialDecryption(0, 0), R.PartialDecryption(0, 0)])


def test_update_lambda_euclidean_div():
    # thresholdkey_test.go:168-177: Div(11 * -7, 3 - 7) = 20 (truncation would give 19)
    assert R.update_lambda(R.PartialDecryption(3, 5), R.PartialDecryption(7, 3), 11) == 20


def test_update_cprime():
    # thresholdkey_test.go:179-190
    tk = R.ThresholdPublicKey(N=99)
    assert R.update_cprime(tk, 77, 52, R.PartialDecryption(3, 5)) == 8558


def test_combine_literal_shares():
    # thresholdkey_test.go:267-281 (TestDecryption)
    tk = R.ThresholdPublicKey(N=637753, Threshold=2, TotalNumberOfDecryptionServers=2, VerificationKey=70661107826)
    shares = [R.PartialDecryption(1, 384111638639), R.PartialDecryption(2, 235243761043)]
    assert R.combine_partial_decryptions(tk, shares) == 100


def test_L():
    # paillier_test.go:20-27
    assert R.L(21, 3) == 6


def test_init_shortcuts():
    # thresholdkey_generator_test.go:213-230
    g = R.ThresholdKeyGenerator(0, 0, p=839, p1=419, q=887, q1=443)
    g.init_shortcuts()
    assert g.n == 744193 and g.m == 185617 and g.nm == 744193 * 185617 and g.n2 == 744193 ** 2


def test_init_d():
    # thresholdkey_generator_test.go:232-243
    g = R.ThresholdKeyGenerator(0, 0, p=863, p1=431, q=839, q1=419)
    g.init_shortcuts()
    g.init_d()
    assert g.d % g.m == 0 and g.d % g.n == 1


def test_are_ps_and_qs_good():
    # thresholdkey_generator_test.go:189-211
    assert not R.ThresholdKeyGenerator(0, 0, p=887, p1=443, q=887, q1=443).are_ps_and_qs_good()
    assert not R.ThresholdKeyGenerator(0, 0, p=887, p1=443, q=443, q1=221).are_ps_and_qs_good()
    assert R.ThresholdKeyGenerator(0, 0, p=887, p1=443, q=839, q1=419).are_ps_and_qs_good()


def test_compute_share():
    # thresholdkey_generator_test.go:282-294
    g = R.ThresholdKeyGenerator(5, 3)
    g.nm = 103
    g.polynomialCoefficients = [29, 88, 51]
    assert g.compute_share(2) == 31


def test_create_verification_keys():
    # thresholdkey_generator_test.go:314-324
    g = R.ThresholdKeyGenerator(10, 0)
    g.v = 54
    g.n2 = 101 * 101
    assert g.create_verification_keys([12, 90, 103]) == [6162, 304, 2728]


# ---- self-consistency at the sizes the reference's randomised tests use -----------------

# two 32-bit primes = 3 mod 4 (paillier.go:131-137) -> the 64-bit keys of the reference's randomised tests
P32, Q32 = 4294967279, 4294967231


def test_round_trip_level1_and_2():
    # paillier_test.go:52-90
    sk, pk = R.keygen_from_primes(P32, Q32)
    for m in (0, 1, 2, 12345, pk.N - 1):
        for level in (R.ENC_LEVEL_ONE, R.ENC_LEVEL_TWO):
            _, ns, _ = pk.moduli_for_level(level)
            mm = m if level == R.ENC_LEVEL_ONE else (m * 7919) % ns
            ct = R.encrypt_with_r_at_level(pk, mm, 987654321, level)
            assert R.decrypt(sk, ct) == mm


def test_homomorphic_ops():
    # operations_test.go:11-54
    sk, pk = R.keygen_from_primes(P32, Q32)
    c1 = R.encrypt_with_r(pk, 13, 1111)
    c2 = R.encrypt_with_r(pk, 19, 2222)
    assert R.decrypt(sk, R.add(pk, c1, c2)) == 32
    assert R.decrypt(sk, R.sub(pk, c2, c1)) == 6
    assert R.decrypt(sk, R.const_mult(pk, c1, 10)) == 130
    assert R.const_mult(pk, c1, 0).C == 1 and R.const_mult(pk, c1, -5).C == 1


def test_extract_randomness():
    # operations_test.go:130-163 (r = i^2 with EncryptWithRAtLevel)
    sk, pk = R.keygen_from_primes(P32, Q32)
    for i in range(2, 20):
        ct = R.encrypt_with_r_at_level(pk, 5 * i, i * i, R.ENC_LEVEL_ONE)
        assert R.extract_randomness(sk, ct) == i * i


def test_threshold_end_to_end_and_zkp():
    # thresholdkey_test.go:192-265,283-292,329-355
    keys = R.threshold_keys_from(p=839, q=887, l=10, w=6, v_seed=123457, coeffs=[11, 22, 33, 44, 55])
    pk = R.PublicKey(N=keys[0].N)
    c = R.encrypt_with_r(pk, 876, 4321).C
    shares = [R.partial_decrypt(k, c) for k in keys]
    tk = R.threshold_public_key(keys[0])
    assert R.combine_partial_decryptions(tk, shares) == 876
    assert R.combine_partial_decryptions(tk, shares[2:8]) == 876
    assert R.combine_partial_decryptions(tk, [shares[i] for i in (9, 0, 3, 5, 7, 2)]) == 876
    with pytest.raises(R.ThresholdError):
        R.combine_partial_decryptions(tk, shares[:5])
    zk = R.partial_decryption_with_zkp(keys[6], c, r=987654321)
    assert R.verify_proof(zk)
    zk.ID += 1                      # wrong verification key -> reject (thresholdkey_test.go:283-292)
    assert not R.verify_proof(zk)


def test_ddleq_completeness_and_soundness():
    # ddleq_test.go:9-72
    sk, pk = R.keygen_from_primes(P32, Q32)
    inner = R.encrypt_with_r(pk, 77, 1234567)
    ct1 = R.encrypt_with_r_at_level(pk, inner.C, 7654321, R.ENC_LEVEL_TWO)
    a, b = 1122334455, 998877665
    ct2 = R.nested_randomize_with(pk, ct1, a, b)
    assert R.nested_decrypt(sk, ct2) == 77
    xs = [1000003 + 17 * i for i in range(10)]
    ys = [2000003 + 29 * i for i in range(10)]
    proof = R.prove_ddleq(sk, 10, ct1, ct2, a, b, xs, ys)
    assert R.verify_ddleq(pk, ct1, ct2, proof)
    bits = {R.random_oracle_bit(ct1.C, ct2.C, p.X, p.Y, p.Alpha) for p in proof}
    assert bits == {True, False}
    other = R.encrypt_with_r_at_level(pk, inner.C + 1, 7654321, R.ENC_LEVEL_TWO)
    assert not R.verify_ddleq(pk, other, ct2, proof)


def test_random_oracle_skips_first_argument():
    # random_oracle.go:24-26
    assert R.random_oracle_digest(1, 2, 3) == R.random_oracle_digest(99, 2, 3)
    import hashlib
    assert R.random_oracle_digest(0, 0x0102, 0) == hashlib.sha256(b"\x01\x02").digest()


def test_safe_prime_candidate_shape():
    # safe_prime_test.go:11-67 shape checks via utils_test.go:66-82
    import random
    rnd = random.Random(7)
    found = 0
    for _ in range(4000):
        raw = bytes(rnd.getrandbits(8) for _ in range((31 + 7) // 8))
        p, q, ok = R.safe_prime_candidate(raw, 32)
        if ok:
            found += 1
            assert p == 2 * q + 1 and p.bit_length() == 32 and q.bit_length() == 31
            assert R._is_probable_prime(p, 20) and R._is_probable_prime(q, 20)
    assert found > 0
