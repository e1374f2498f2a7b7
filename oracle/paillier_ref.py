"""CPU oracle: a function-by-function restatement of sachaservan/paillier.

TEST INFRASTRUCTURE ONLY.  Nothing under paillier_b200/ may import this module;
it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
as the checker, never as the thing measured or shipped.

Every function cites the reference lines it follows.  Arithmetic is Python's
arbitrary-precision int (pow/divmod are C), standing in for github.com/ncw/gmp
(un-vendored, version unpinned in the reference: there is no go.mod), whose
semantics are mirrored where they differ from the obvious:

  * gmp.Int.Exp(x, y, m): y <= 0 -> 1; m == nil -> plain power (mpz_pow_ui);
    else mpz_powm with the result in [0, |m|).
  * Div / Mod are Euclidean (thresholdkey_test.go:168-177 pins Div(-77, -4) = 20).
  * Bytes() is the minimal big-endian magnitude, 0 -> b"".

Parity pinning: checked against every known-answer test the reference holds
(tests/test_oracle_kats.py lists them with file:line).  The reference has no
golden vectors at 1024-bit and above and cannot be built here (no Go toolchain,
ncw/gmp absent): PARITY UNPINNED at >= 1024 bits against the reference itself.
There it rests on three unrelated bignum implementations agreeing on the same
inputs -- this module (CPython ints), libgmp (oracle/gmp_ref.c) and OpenSSL BN
(oracle/ossl_ref.c) -- see DESIGN.md section 2.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------
# ncw/gmp semantics
# ----------------------------------------------------------------------------

def gmp_exp(x: int, y: int, m: Optional[int]) -> int:
    """gmp.Int.Exp as used at every Exp call site (census in SURVEY.md 8b)."""
    if y <= 0:
        return 1
    if m is None or m == 0:
        return x ** y
    return pow(x, y, abs(m))


def gmp_div(x: int, y: int) -> int:
    """Euclidean division (Go big.Int.Div semantics kept by ncw/gmp)."""
    q, r = divmod(x, y)          # floor division
    if r < 0:                    # only when y < 0
        q += 1
    return q


def gmp_mod(x: int, y: int) -> int:
    return x - gmp_div(x, y) * y


def gmp_mod_inverse(a: int, m: int) -> int:
    """mpz_invert: result in [0, m); undefined for non-units (we raise)."""
    return pow(a, -1, m)


def gmp_bytes(x: int) -> bytes:
    x = abs(x)
    return x.to_bytes((x.bit_length() + 7) // 8, "big")


# ----------------------------------------------------------------------------
# utils.go
# ----------------------------------------------------------------------------

def factorial(n: int) -> int:
    """utils.go:17-23"""
    ret = 1
    for i in range(1, n + 1):
        ret *= i
    return ret


def L(u: int, n: int) -> int:
    """paillier.go:437-440"""
    return gmp_div(u - 1, n)


# ----------------------------------------------------------------------------
# paillier.go
# ----------------------------------------------------------------------------
ENC_LEVEL_ONE = 0
ENC_LEVEL_TWO = 1
REGULAR, ALTERNATIVE, MIXED = 0, 1, 2


@dataclass
class PublicKey:
    """paillier.go:46-56"""
    N: int
    G: int = 0
    H: int = 0
    K: int = 0

    def __post_init__(self):
        if not self.G:
            self.G = self.N + 1                      # paillier.go:147

    def n2(self) -> int:                             # paillier.go:72-79
        return self.N * self.N

    def n3(self) -> int:                             # paillier.go:82-90
        return self.N * self.N * self.N

    def moduli_for_level(self, level: int) -> Tuple[int, int, int]:
        """paillier.go:403-414 -> (s, n^s, n^(s+1))"""
        if level == ENC_LEVEL_TWO:
            return 2, self.n2(), self.n3()
        return 1, self.N, self.n2()

    def qr_generator_for_level(self, level: int) -> int:
        """paillier.go:416-434"""
        if level == ENC_LEVEL_ONE:
            return gmp_exp(self.N - self.H, self.N, self.n2())
        return gmp_exp(self.n2() - self.H, self.n2(), self.n3())


@dataclass
class SecretKey(PublicKey):
    """paillier.go:59-62; Lambda = (p-1)(q-1) (paillier.go:152,452-454)"""
    Lambda: int = 0


@dataclass
class Ciphertext:
    """paillier.go:65-69"""
    C: int
    Level: int = ENC_LEVEL_ONE
    EncMethod: int = REGULAR


def keygen_from_primes(p: int, q: int, h_seed_r: Optional[int] = None) -> Tuple[SecretKey, PublicKey]:
    """paillier.go:106-179 with the primes supplied (the reference draws them from crypto/rand)."""
    assert p != q and p % 4 == 3 and q % 4 == 3
    n = p * q
    bits = n.bit_length()
    k = 2 ** (bits // 2)                             # paillier.go:151 (secparam/2)
    lam = (p - 1) * (q - 1)                          # paillier.go:152
    h = (h_seed_r * h_seed_r) % n if h_seed_r else 0  # utils.go:53-59
    pk = PublicKey(N=n, G=n + 1, H=h, K=k)
    sk = SecretKey(N=n, G=n + 1, H=h, K=k, Lambda=lam)
    return sk, pk


def encrypt_with_r_at_level(pk: PublicKey, m: int, r: int, level: int = ENC_LEVEL_ONE) -> Ciphertext:
    """paillier.go:206-218"""
    _, ns, ns1 = pk.moduli_for_level(level)
    gm = gmp_exp(pk.G, m, ns1)
    rn = gmp_exp(r, ns, ns1)
    return Ciphertext(gmp_mod(gm * rn, ns1), level, REGULAR)


def encrypt_with_r(pk: PublicKey, m: int, r: int) -> Ciphertext:
    """paillier.go:185-187"""
    return encrypt_with_r_at_level(pk, m, r, ENC_LEVEL_ONE)


def alt_encrypt_with_r_at_level(pk: PublicKey, m: int, r: int, level: int = ENC_LEVEL_ONE) -> Tuple[Ciphertext, int]:
    """paillier.go:221-238; returns the ciphertext and the caller's r as mutated in place (:228)."""
    _, _, ns1 = pk.moduli_for_level(level)
    h = pk.qr_generator_for_level(level)
    r = gmp_mod(r, pk.K)
    gm = gmp_exp(pk.G, m, ns1)
    hr = gmp_exp(h, r, ns1)
    return Ciphertext(gmp_mod(gm * hr, ns1), level, ALTERNATIVE), r


def recovery_algorithm(sk: SecretKey, a: int, s: int) -> int:
    """paillier.go:308-340 (Damgard-Jurik recovery of m*lambda mod n^s)"""
    i = 0
    for j in range(1, s + 1):
        nj = gmp_exp(sk.N, j, None)
        nj1 = gmp_exp(sk.N, j + 1, None)
        amod = gmp_mod(a, nj1)
        t1 = L(amod, sk.N)
        t2 = i
        for k in range(2, j + 1):
            nk = gmp_exp(sk.N, k - 1, None)
            i = i - 1
            t2 = gmp_mod(t2 * i, nj)
            t2 = t2 * nk
            kfac = gmp_mod_inverse(factorial(k), nj)
            t2 = t2 * kfac
            t2 = t1 - t2
            t1 = gmp_mod(t2, nj)
        i = t1
    return i


def decrypt(sk: SecretKey, ct: Ciphertext) -> int:
    """paillier.go:292-303"""
    s, ns, ns1 = sk.moduli_for_level(ct.Level)
    tmp = gmp_exp(ct.C, sk.Lambda, ns1)
    ml = recovery_algorithm(sk, tmp, s)
    mu = gmp_mod_inverse(sk.Lambda, ns)
    return gmp_mod(ml * mu, ns)


def nested_decrypt(sk: SecretKey, ct: Ciphertext) -> int:
    """paillier.go:344-372"""
    assert ct.Level == ENC_LEVEL_TWO, "no nested ciphertexts to recover"
    ct1 = Ciphertext(decrypt(sk, ct), ENC_LEVEL_ONE, MIXED)
    if ct1.C == 0:
        return 0
    return decrypt(sk, ct1)


# ----------------------------------------------------------------------------
# operations.go
# ----------------------------------------------------------------------------

def add(pk: PublicKey, *cts: Ciphertext) -> Ciphertext:
    """operations.go:11-29"""
    acc = 1
    level = cts[0].Level
    _, _, ns1 = pk.moduli_for_level(level)
    for c in cts:
        acc = gmp_mod(acc * c.C, ns1)
    return Ciphertext(acc, level, MIXED)


def sub(pk: PublicKey, *cts: Ciphertext) -> Ciphertext:
    """operations.go:32-55 (one argument: the input value, unreduced, :34)"""
    acc = cts[0].C
    level = cts[0].Level
    _, _, ns1 = pk.moduli_for_level(level)
    for c in cts[1:]:
        acc = gmp_mod(acc * gmp_mod_inverse(c.C, ns1), ns1)
    return Ciphertext(acc, level, MIXED)


def const_mult(pk: PublicKey, ct: Ciphertext, k: int) -> Ciphertext:
    """operations.go:58-64 (k <= 0 -> 1 through gmp Exp)"""
    _, _, ns1 = pk.moduli_for_level(ct.Level)
    return Ciphertext(gmp_exp(ct.C, k, ns1), ct.Level, ct.EncMethod)


def extract_randomness(sk: SecretKey, ct: Ciphertext) -> int:
    """operations.go:75-91"""
    _, ns, ns1 = sk.moduli_for_level(ct.Level)
    ns_inv = gmp_mod_inverse(ns, sk.Lambda)
    v = decrypt(sk, ct)
    gv_inv = gmp_mod_inverse(gmp_exp(sk.G, v, ns1), ns1)
    z = gmp_mod(gv_inv * ct.C, ns1)
    return gmp_exp(z, ns_inv, sk.N)


def nested_randomize_with(pk: PublicKey, ct: Ciphertext, a: int, b: int) -> Ciphertext:
    """operations.go:96-118 with the randomness (a, b) supplied"""
    assert ct.Level == ENC_LEVEL_TWO
    n, n2, n3 = pk.N, pk.n2(), pk.n3()
    an = gmp_exp(a, n, n2)
    bn2 = gmp_exp(b, n2, n3)
    r = gmp_mod(gmp_exp(ct.C, an, n3) * bn2, n3)
    return Ciphertext(r, ct.Level, REGULAR)


def nested_add(pk: PublicKey, ct1: Ciphertext, ct2: Ciphertext) -> Ciphertext:
    """operations.go:121-127"""
    assert ct1.Level == ENC_LEVEL_TWO and ct2.Level == ENC_LEVEL_ONE
    return const_mult(pk, ct1, ct2.C)


def nested_sub(pk: PublicKey, ct1: Ciphertext, ct2: Ciphertext) -> Ciphertext:
    """operations.go:130-140"""
    assert ct1.Level == ENC_LEVEL_TWO and ct2.Level == ENC_LEVEL_ONE
    _, _, ns1 = pk.moduli_for_level(ct2.Level)
    return const_mult(pk, ct1, gmp_mod_inverse(ct2.C, ns1))


# ----------------------------------------------------------------------------
# random_oracle.go
# ----------------------------------------------------------------------------

def random_oracle_digest(*values: int) -> bytes:
    """random_oracle.go:20-32 -- the FIRST argument is skipped (:24-26)"""
    data = b"".join(gmp_bytes(v) for v in values[1:])
    return hashlib.sha256(data).digest()


def random_oracle_bit(*values: int) -> bool:
    """random_oracle.go:10-17"""
    return int.from_bytes(random_oracle_digest(*values), "big") % 2 == 1


# ----------------------------------------------------------------------------
# thresholdkey.go
# ----------------------------------------------------------------------------

@dataclass
class ThresholdPublicKey:
    """thresholdkey.go:26-32"""
    N: int
    TotalNumberOfDecryptionServers: int = 0
    Threshold: int = 0
    VerificationKey: int = 0
    VerificationKeys: List[int] = field(default_factory=list)

    def n2(self) -> int:
        return self.N * self.N

    def delta(self) -> int:
        """thresholdkey.go:70-72"""
        return factorial(self.TotalNumberOfDecryptionServers)

    def combine_shares_constant(self) -> int:
        """thresholdkey.go:63-66"""
        return gmp_mod_inverse(4 * self.delta() * self.delta(), self.N)


@dataclass
class ThresholdSecretKey(ThresholdPublicKey):
    """thresholdkey.go:38-42"""
    ID: int = 0
    Share: int = 0


@dataclass
class PartialDecryption:
    """thresholdkey.go:45-48"""
    ID: int
    Decryption: int


@dataclass
class PartialDecryptionZKP:
    """thresholdkey.go:52-58"""
    ID: int
    Decryption: int
    Key: ThresholdPublicKey
    E: int
    Z: int
    C: int


class ThresholdError(Exception):
    pass


def verify_partial_decryptions(tk: ThresholdPublicKey, shares: Sequence[PartialDecryption]) -> None:
    """thresholdkey.go:77-89"""
    if len(shares) < tk.Threshold:
        raise ThresholdError("Threshold not meet")
    if len({s.ID for s in shares}) != len(shares):
        raise ThresholdError("two shares has been created by the same server")


def update_lambda(share1: PartialDecryption, share2: PartialDecryption, lam: int) -> int:
    """thresholdkey.go:91-95"""
    return gmp_div(lam * (-share2.ID), share1.ID - share2.ID)


def compute_lambda(tk: ThresholdPublicKey, share: PartialDecryption, shares: Sequence[PartialDecryption]) -> int:
    """thresholdkey.go:99-107"""
    lam = tk.delta()
    for share2 in shares:
        if share2.ID != share.ID:
            lam = update_lambda(share, share2, lam)
    return lam


def tk_exp(a: int, b: int, c: int) -> int:
    """thresholdkey.go:132-138"""
    if b < 0:
        return gmp_mod_inverse(gmp_exp(a, -b, c), c)
    return gmp_exp(a, b, c)


def update_cprime(tk: ThresholdPublicKey, cprime: int, lam: int, share: PartialDecryption) -> int:
    """thresholdkey.go:119-124"""
    ret = tk_exp(share.Decryption, 2 * lam, tk.n2())
    return gmp_mod(cprime * ret, tk.n2())


def compute_decryption(tk: ThresholdPublicKey, cprime: int) -> int:
    """thresholdkey.go:143-146"""
    return gmp_mod(tk.combine_shares_constant() * L(cprime, tk.N), tk.N)


def combine_partial_decryptions(tk: ThresholdPublicKey, shares: Sequence[PartialDecryption]) -> int:
    """thresholdkey.go:149-161"""
    verify_partial_decryptions(tk, shares)
    cprime = 1
    for share in shares:
        lam = compute_lambda(tk, share, shares)
        cprime = update_cprime(tk, cprime, lam, share)
    return compute_decryption(tk, cprime)


def partial_decrypt(tsk: ThresholdSecretKey, c: int) -> PartialDecryption:
    """thresholdkey.go:192-201"""
    exp = tsk.Share * (2 * tsk.delta())
    return PartialDecryption(tsk.ID, gmp_exp(c, exp, tsk.n2()))


def threshold_public_key(tsk: ThresholdSecretKey) -> ThresholdPublicKey:
    """thresholdkey.go:213-221"""
    return ThresholdPublicKey(N=tsk.N, TotalNumberOfDecryptionServers=tsk.TotalNumberOfDecryptionServers,
                              Threshold=tsk.Threshold, VerificationKey=tsk.VerificationKey,
                              VerificationKeys=list(tsk.VerificationKeys))


def zkp_hash(a: int, b: int, c4: int, ci2: int) -> int:
    """thresholdkey.go:319-326"""
    h = hashlib.sha256()
    for v in (a, b, c4, ci2):
        h.update(gmp_bytes(v))
    return int.from_bytes(h.digest(), "big")


def partial_decryption_with_zkp(tsk: ThresholdSecretKey, c: int, r: int) -> PartialDecryptionZKP:
    """thresholdkey.go:225-255 with the random r in [0, n^2) supplied (the reference draws it at :233)"""
    dec = partial_decrypt(tsk, c).Decryption
    n2 = tsk.n2()
    c4 = gmp_exp(c, 4, None)                       # unreduced, :241
    a = gmp_exp(c4, r, n2)
    b = gmp_exp(tsk.VerificationKey, r, n2)
    ci2 = gmp_exp(dec, 2, None)                    # unreduced, :248
    e = zkp_hash(a, b, c4, ci2)
    z = r + e * tsk.delta() * tsk.Share            # :313-317
    return PartialDecryptionZKP(tsk.ID, dec, threshold_public_key(tsk), e, z, c)


def verify_part1(pd: PartialDecryptionZKP) -> int:
    """thresholdkey.go:293-302"""
    n2 = pd.Key.n2()
    c4 = gmp_exp(pd.C, 4, None)
    dec2 = gmp_exp(pd.Decryption, 2, None)
    a1 = gmp_exp(c4, pd.Z, n2)
    a2 = gmp_mod_inverse(gmp_exp(dec2, pd.E, n2), n2)
    return gmp_mod(a1 * a2, n2)


def verify_part2(pd: PartialDecryptionZKP) -> int:
    """thresholdkey.go:304-311"""
    n2 = pd.Key.n2()
    vi = pd.Key.VerificationKeys[pd.ID - 1]
    b1 = gmp_exp(pd.Key.VerificationKey, pd.Z, n2)
    b2 = gmp_mod_inverse(gmp_exp(vi, pd.E, n2), n2)
    return gmp_mod(b1 * b2, n2)


def verify_proof(pd: PartialDecryptionZKP) -> bool:
    """thresholdkey.go:278-291"""
    a = verify_part1(pd)
    b = verify_part2(pd)
    c4 = gmp_exp(pd.C, 4, None)
    ci2 = gmp_exp(pd.Decryption, 2, None)
    return pd.E == zkp_hash(a, b, c4, ci2)


def combine_partial_decryptions_zkp(tk: ThresholdPublicKey, shares: Sequence[PartialDecryptionZKP]) -> int:
    """thresholdkey.go:164-172"""
    ok = [PartialDecryption(s.ID, s.Decryption) for s in shares if verify_proof(s)]
    return combine_partial_decryptions(tk, ok)


# ----------------------------------------------------------------------------
# thresholdkey_generator.go (with the primes and random draws supplied)
# ----------------------------------------------------------------------------

@dataclass
class ThresholdKeyGenerator:
    """thresholdkey_generator.go:19-44"""
    TotalNumberOfDecryptionServers: int
    Threshold: int
    p: int = 0
    p1: int = 0
    q: int = 0
    q1: int = 0
    n: int = 0
    m: int = 0
    n2: int = 0
    nm: int = 0
    d: int = 0
    v: int = 0
    polynomialCoefficients: List[int] = field(default_factory=list)

    def are_ps_and_qs_good(self) -> bool:
        """thresholdkey_generator.go:120-131"""
        return not (self.p == self.q or self.p == self.q1 or self.p1 == self.q)

    def init_shortcuts(self) -> None:
        """thresholdkey_generator.go:113-118"""
        self.n = self.p * self.q
        self.m = self.p1 * self.q1
        self.n2 = self.n * self.n
        self.nm = self.n * self.m

    def init_d(self) -> None:
        """thresholdkey_generator.go:177-180"""
        self.d = gmp_mod_inverse(self.m, self.n) * self.m

    def compute_v(self, r: int) -> None:
        """thresholdkey_generator.go:147-151 + utils.go:53-59 with r in Z*_{n^2} supplied"""
        self.v = gmp_mod(r * r, self.n2)

    def set_polynomial(self, coeffs: Sequence[int]) -> None:
        """thresholdkey_generator.go:197-209: a_0 = d, a_i uniform in [0, nm)"""
        self.polynomialCoefficients = [self.d] + list(coeffs)
        assert len(self.polynomialCoefficients) == self.Threshold

    def compute_share(self, index: int) -> int:
        """thresholdkey_generator.go:213-223"""
        share = 0
        for i in range(self.Threshold):
            share += self.polynomialCoefficients[i] * gmp_exp(index + 1, i, None)
        return gmp_mod(share, self.nm)

    def create_shares(self) -> List[int]:
        """thresholdkey_generator.go:225-231"""
        return [self.compute_share(i) for i in range(self.TotalNumberOfDecryptionServers)]

    def delta(self) -> int:
        """thresholdkey_generator.go:233-235"""
        return factorial(self.TotalNumberOfDecryptionServers)

    def create_verification_keys(self, shares: Sequence[int]) -> List[int]:
        """thresholdkey_generator.go:246-254"""
        d = self.delta()
        return [gmp_exp(self.v, s * d, self.n2) for s in shares]

    def create_private_keys(self) -> List[ThresholdSecretKey]:
        """thresholdkey_generator.go:256-278"""
        shares = self.create_shares()
        vks = self.create_verification_keys(shares)
        return [ThresholdSecretKey(N=self.n, TotalNumberOfDecryptionServers=self.TotalNumberOfDecryptionServers,
                                   Threshold=self.Threshold, VerificationKey=self.v, VerificationKeys=vks,
                                   ID=i + 1, Share=shares[i])
                for i in range(self.TotalNumberOfDecryptionServers)]


def threshold_keys_from(p: int, q: int, l: int, w: int, v_seed: int, coeffs: Sequence[int]) -> List[ThresholdSecretKey]:
    """GenerateKeys (thresholdkey_generator.go:47-55) with safe primes p = 2p1+1, q = 2q1+1 and all draws supplied."""
    g = ThresholdKeyGenerator(l, w, p=p, p1=(p - 1) // 2, q=q, q1=(q - 1) // 2)
    assert g.are_ps_and_qs_good()
    g.init_shortcuts()
    g.init_d()
    g.compute_v(v_seed)
    g.set_polynomial(coeffs)
    return g.create_private_keys()


# ----------------------------------------------------------------------------
# ddleq.go (with the per-instance randomness supplied)
# ----------------------------------------------------------------------------

@dataclass
class DDLEQProofInstance:
    """ddleq.go:11-13"""
    X: int
    Y: int
    Alpha: int
    E: int
    F: int


def prove_ddleq_instance(sk: SecretKey, ct1: Ciphertext, ct2: Ciphertext, a: int, b: int, x: int, y: int) -> DDLEQProofInstance:
    """ddleq.go:55-127 with x, y in Z*_n supplied (drawn at :71-79 in the reference)"""
    n, n2, n3 = sk.N, sk.n2(), sk.n3()
    sanity = gmp_mod(gmp_exp(ct1.C, gmp_exp(a, n, n2), n3) * gmp_exp(b, n2, n3), n3)
    if sanity != ct2.C:
        raise ValueError("cannot prove re-encryption because inputs are wrong")
    xn = gmp_exp(x, n, n2)
    yn2 = gmp_exp(y, n2, n3)
    alpha = gmp_mod(gmp_exp(ct1.C, xn, n3) * yn2, n3)
    chal = random_oracle_bit(ct1.C, ct2.C, x, y, alpha)
    e = x
    if chal:
        e = gmp_mod(e * gmp_mod_inverse(a, n2), n2)
    f = y
    if chal:
        s = extract_randomness(sk, ct1)
        an = gmp_exp(a, n, n2)
        en = gmp_exp(e, n, n2)
        c = gmp_exp(s, an, n3)
        c = c * b
        c = gmp_exp(c, en, n3)
        c = gmp_mod_inverse(c, n3)
        c = c * gmp_exp(s, xn, n3)
        f = gmp_mod(f * c, n3)
    return DDLEQProofInstance(x, y, alpha, e, f)


def prove_ddleq(sk: SecretKey, secpar: int, ct1: Ciphertext, ct2: Ciphertext, a: int, b: int,
                xs: Sequence[int], ys: Sequence[int]) -> List[DDLEQProofInstance]:
    """ddleq.go:27-40"""
    return [prove_ddleq_instance(sk, ct1, ct2, a, b, xs[i], ys[i]) for i in range(secpar)]


def verify_ddleq_instance(pk: PublicKey, ct1: Ciphertext, ct2: Ciphertext, proof: DDLEQProofInstance) -> bool:
    """ddleq.go:129-153"""
    n, n2, n3 = pk.N, pk.n2(), pk.n3()
    chal = random_oracle_bit(ct1.C, ct2.C, proof.X, proof.Y, proof.Alpha)
    check = ct2.C if chal else ct1.C
    en = gmp_exp(proof.E, n, n2)
    fn2 = gmp_exp(proof.F, n2, n3)
    check = gmp_mod(gmp_exp(check, en, n3) * fn2, n3)
    return proof.Alpha == check


def verify_ddleq(pk: PublicKey, ct1: Ciphertext, ct2: Ciphertext, proof: Sequence[DDLEQProofInstance]) -> bool:
    """ddleq.go:44-53"""
    return all(verify_ddleq_instance(pk, ct1, ct2, inst) for inst in proof)


# ----------------------------------------------------------------------------
# safe_prime.go (candidate procedure on a supplied byte string)
# ----------------------------------------------------------------------------
SMALL_PRIMES = [3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53]   # safe_prime.go:26-28
SMALL_PRIMES_PRODUCT = 16294579238595022365                                # safe_prime.go:34


def is_prime_candidate(number: int) -> bool:
    """safe_prime.go:280-290"""
    m = number % SMALL_PRIMES_PRODUCT
    return not any(m % p == 0 and m != p for p in SMALL_PRIMES)


def is_pocklington_criterion_satisfied(p: int) -> bool:
    """safe_prime.go:272-278"""
    return pow(2, p - 1, p) == 1


def _is_probable_prime(n: int, rounds: int = 20) -> bool:
    """Stand-in for Go's big.Int.ProbablyPrime(20) (safe_prime.go:256): decisions agree for every
    input except with negligible probability (true primes always pass; composites are rejected by
    Miller-Rabin with fixed small bases + a strong Lucas-free deterministic fallback)."""
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    bases = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71][:max(rounds, 1)]
    for a in bases:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def safe_prime_candidate(raw: bytes, p_bit_len: int) -> Tuple[int, int, bool]:
    """One iteration of runGenPrimeRoutine's loop (safe_prime.go:170-263) on the byte string the
    reference would have read from its io.Reader.  Returns (p, q, accepted)."""
    q_bit_len = p_bit_len - 1
    b = q_bit_len % 8 or 8
    bs = bytearray(raw)
    assert len(bs) == (q_bit_len + 7) // 8
    bs[0] &= (1 << b) - 1
    if b >= 2:
        bs[0] |= 3 << (b - 2)
    else:
        bs[0] |= 1
        if len(bs) > 1:
            bs[1] |= 0x80
    bs[-1] |= 1
    q = int.from_bytes(bs, "big")
    mod = q % SMALL_PRIMES_PRODUCT
    p = 0                                           # reference: p keeps its previous value if no delta survives
    delta = 0
    while delta < (1 << 20):
        m = mod + delta
        if any(m % pr == 0 and (q_bit_len > 6 or m != pr) for pr in SMALL_PRIMES):
            delta += 2
            continue
        # safe_prime.go:221-224: q is advanced by this delta (cumulatively) before the remaining filters
        if delta > 0:
            q += delta
        if q % 3 == 1:
            delta += 2
            continue
        p = 2 * q + 1
        if not is_prime_candidate(p):
            delta += 2
            continue
        break
    ok = _is_probable_prime(q, 20) and is_pocklington_criterion_satisfied(p) and q.bit_length() == q_bit_len
    return p, q, ok
