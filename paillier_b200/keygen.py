"""Host-side mirror of thresholdkey_generator.go (row a18 of SURVEY.md 8a).

The key-generation glue is a handful of big-integer operations per key, done once on the host
(thresholdkey_generator.go:113-231 are O(l*w) multiplications); the l modular exponentiations of
createVerificationKeys (thresholdkey_generator.go:246-254) go through the GPU engine's batched Exp.
Randomness defaults to the operating system's CSPRNG (`random.SystemRandom`, i.e. os.urandom), as the reference draws
every prime, coefficient and v from crypto/rand (thresholdkey_generator.go:66, paillier.go:122-139).  A seeded
`random.Random` may be passed as `rng` for reproducible TESTS ONLY: keys made from it are predictable.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

from .api import PublicKey, ThresholdSecretKey


def factorial(n: int) -> int:
    """utils.go:17-23"""
    r = 1
    for i in range(2, n + 1):
        r *= i
    return r


@dataclass
class ThresholdKeyGenerator:
    """thresholdkey_generator.go:19-44, 62-86"""
    PublicKeyBitLength: int
    TotalNumberOfDecryptionServers: int
    Threshold: int
    rng: random.Random = field(default_factory=random.SystemRandom)      # CSPRNG; a seeded random.Random is for tests only
    p: int = 0
    q: int = 0
    batch: int = 1 << 15          # safe-prime candidates per GPU call
    max_batches: int = 4096

    def __post_init__(self):
        if self.PublicKeyBitLength % 2 == 1:
            raise ValueError("Public key bit length must be an even number")      # :68-73
        if self.PublicKeyBitLength < 18:
            raise ValueError("Public key bit length must be at least 18 bits")    # :74-78

    def with_safe_primes(self, p: int, q: int) -> "ThresholdKeyGenerator":
        """Supply p = 2p1+1, q = 2q1+1 (the reference searches for them, :88-99)."""
        self.p, self.q = p, q
        return self

    def generateSafePrimes(self, device: int = 0):
        """thresholdkey_generator.go:88-99: a safe prime of PublicKeyBitLength/2 bits, candidates tested on the GPU in
        stream order of this generator's randomness"""
        reader = lambda nbytes: self.rng.randbytes(nbytes)
        return GenerateSafePrime(self.PublicKeyBitLength // 2, reader, batch=self.batch, max_batches=self.max_batches, device=device)

    def initPsAndQs(self, device: int = 0) -> None:
        """thresholdkey_generator.go:133-144 (retry until p, q, p1, q1 are pairwise usable, :120-131)"""
        while True:
            self.p, _ = self.generateSafePrimes(device)
            self.q, _ = self.generateSafePrimes(device)
            p1, q1 = (self.p - 1) // 2, (self.q - 1) // 2
            if self.p != self.q and self.p != q1 and p1 != self.q:
                return

    def GenerateKeys(self, device: int = 0) -> List[ThresholdSecretKey]:
        """thresholdkey_generator.go:47-55"""
        if not self.p or not self.q:
            self.initPsAndQs(device)
        p, q = self.p, self.q
        p1, q1 = (p - 1) // 2, (q - 1) // 2
        if p == q or p == q1 or p1 == q:                                           # :120-131
            raise ValueError("bad safe primes")
        n, m = p * q, p1 * q1                                                      # :113-118
        n2, nm = n * n, n * m
        d = pow(m, -1, n) * m                                                      # :177-180
        while True:                                                                # utils.go:36-59 over n^2
            r = self.rng.randrange(1, n2)
            if r % p and r % q:
                break
        v = r * r % n2                                                             # :147-151
        coeffs = [d] + [self.rng.randrange(nm) for _ in range(1, self.Threshold)]  # :197-209
        l = self.TotalNumberOfDecryptionServers
        shares = [sum(a * (i + 1) ** k for k, a in enumerate(coeffs)) % nm for i in range(l)]   # :213-231
        delta = factorial(l)
        pk = PublicKey(n, device)
        try:
            vks = pk.ExpBatch([v] * l, [s * delta for s in shares])                # :246-254 on the GPU
        finally:
            pk.close()
        return [ThresholdSecretKey(n, l, self.Threshold, v, vks, ID=i + 1, Share=shares[i], device=device)
                for i in range(l)]                                                 # :256-278


# ---- safe_prime.go ---------------------------------------------------------------------------------------

def _prime_shape(p_bits: int) -> int:
    if p_bits <= 1024:
        return 32
    if p_bits <= 1536:
        return 48
    if p_bits <= 2048:
        return 64
    raise ValueError("safe primes of up to 2048 bits are supported")


def safe_prime_scan(p_bit_len: int, raw: bytes, device: int = 0):
    """The candidate procedure of runGenPrimeRoutine (safe_prime.go:170-263) on the GPU for every
    ceil((p_bit_len-1)/8)-byte string of `raw`, in stream order.  Returns (ps, qs, accepted)."""
    import ctypes as C
    import numpy as np
    from ._lib import lib, PGPU_OK, PgpuError
    from .api import from_records
    nb = (p_bit_len - 1 + 7) // 8
    count = len(raw) // nb
    S = _prime_shape(p_bit_len)
    buf = np.frombuffer(raw[:count * nb], dtype=np.uint8).copy()
    p = np.zeros(count * S * 4, dtype=np.uint8)
    q = np.zeros(count * S * 4, dtype=np.uint8)
    ok = np.zeros(count, dtype=np.uint8)
    nl = C.c_uint64()
    rc = lib.pgpu_safe_prime_scan(device, p_bit_len, count, buf.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                                  q.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p), C.byref(nl))
    if rc != PGPU_OK:
        raise PgpuError(rc, (lib.pgpu_primes_last_error() or b"").decode())
    return from_records(p, S * 4), from_records(q, S * 4), [bool(x) for x in ok]


def miller_rabin(bits: int, candidates: Sequence[int], rounds: int = 20, device: int = 0) -> List[bool]:
    """big.Int.ProbablyPrime stand-in on the GPU: `rounds` strong tests (bases 2, 3, 5, ...)"""
    import ctypes as C
    import numpy as np
    from ._lib import lib, PGPU_OK, PgpuError
    from .api import to_records
    S = _prime_shape(bits)
    rec = to_records(candidates, S * 4)
    ok = np.zeros(len(candidates), dtype=np.uint8)
    rc = lib.pgpu_miller_rabin(device, bits, len(candidates), rec.ctypes.data_as(C.c_void_p), rounds,
                               ok.ctypes.data_as(C.c_void_p), None)
    if rc != PGPU_OK:
        raise PgpuError(rc, (lib.pgpu_primes_last_error() or b"").decode())
    return [bool(x) for x in ok]


def GenerateSafePrime(bitLen: int, random_reader, batch: int = 1 << 14, max_batches: int = 64, device: int = 0):
    """safe_prime.go:61-105 with the candidate loop on the GPU: reads `batch` candidates at a time from
    `random_reader(nbytes) -> bytes` and returns (p, q) of the first accepted candidate in stream order.
    The reference's goroutine race (first finisher wins) and its wall-clock timeout become a bound on
    the number of batches."""
    if bitLen < 6:
        raise ValueError("safe prime size must be at least 6 bits")               # :67-69
    nb = (bitLen - 1 + 7) // 8
    for _ in range(max_batches):
        ps, qs, ok = safe_prime_scan(bitLen, random_reader(batch * nb), device)
        for p, q, good in zip(ps, qs, ok):
            if good:
                return p, q
    raise TimeoutError(f"generator gave up after {max_batches} batches of {batch} candidates")   # :101-103


# ---- paillier.go KeyGen ----------------------------------------------------------------------------------

def _random_prime_3mod4(bits: int, rng, device: int, batch: int = 4096) -> int:
    """Stand-in for crypto/rand.Prime(bits) followed by KeyGen's `= 3 (mod 4)` filter (paillier.go:122-139): random odd
    candidates with the two top bits set (as rand.Prime does) and the low two bits set, Miller-Rabin (20 rounds) for the
    whole batch on the GPU, first survivor in stream order wins."""
    while True:
        cands = [rng.getrandbits(bits) | (3 << (bits - 2)) | 3 for _ in range(batch)]
        small = (3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53)
        keep = [c for c in cands if all(c % s for s in small)]
        for c, ok in zip(keep, miller_rabin(bits, keep, 20, device)):
            if ok:
                return c


def KeyGen(secparam: int, rng: Optional[random.Random] = None, device: int = 0):
    """paillier.go:106-179: returns (SecretKey, PublicKey) mirrors bound to `device`; p, q = 3 (mod 4), distinct;
    g = n+1; K = 2^(secparam/2); H = r^2 mod n for a random unit r (utils.go:36-59)."""
    from math import gcd
    from .api import PublicKey as _PK, SecretKey as _SK
    if secparam % 2 != 0:
        raise ValueError("KeyGen: secparam must be divisible by 2")           # :108-110
    if secparam < 64:
        raise ValueError("KeyGen: secparam must be at least 64 bits")         # :112-114
    rng = rng or random.SystemRandom()                                         # crypto/rand, :122-139
    while True:
        p = _random_prime_3mod4(secparam // 2, rng, device)
        q = _random_prime_3mod4(secparam // 2, rng, device)
        if p != q:                                                             # :135-137
            break
    n = p * q
    k = 1 << (secparam // 2)                                                   # :151
    while True:                                                                # utils.go:36-49
        r = rng.randrange(n)
        if r != 0 and gcd(r, n) == 1:
            break
    h = r * r % n                                                              # utils.go:53-59
    sk = _SK(n, p=p, q=q, device=device, H=h, K=k)
    pk = _PK(n, device=device, H=h, K=k)
    return sk, pk
