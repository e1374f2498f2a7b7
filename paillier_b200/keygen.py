"""Host-side mirror of thresholdkey_generator.go (row a18 of SURVEY.md 8a).

The key-generation glue is a handful of big-integer operations per key, done once on the host
(thresholdkey_generator.go:113-231 are O(l*w) multiplications); the l modular exponentiations of
createVerificationKeys (thresholdkey_generator.go:246-254) go through the GPU engine's batched Exp.
Randomness is injected (the reference takes an io.Reader, thresholdkey_generator.go:66).
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
    rng: random.Random = field(default_factory=lambda: random.Random(20260101))
    p: int = 0
    q: int = 0

    def __post_init__(self):
        if self.PublicKeyBitLength % 2 == 1:
            raise ValueError("Public key bit length must be an even number")      # :68-73
        if self.PublicKeyBitLength < 18:
            raise ValueError("Public key bit length must be at least 18 bits")    # :74-78

    def with_safe_primes(self, p: int, q: int) -> "ThresholdKeyGenerator":
        """Supply p = 2p1+1, q = 2q1+1 (the reference searches for them, :88-99)."""
        self.p, self.q = p, q
        return self

    def GenerateKeys(self, device: int = 0) -> List[ThresholdSecretKey]:
        """thresholdkey_generator.go:47-55"""
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
