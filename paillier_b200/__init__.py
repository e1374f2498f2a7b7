"""paillier_b200: B200-native batch engine behind sachaservan/paillier's mod-n^2 exponentiation path.

The product is the C-ABI shared library (include/pgpu.h, paillier_b200/csrc); this package is the
host-side mirror of the reference's interface used by the tests and the benchmark.
"""
from .api import (  # noqa: F401
    Ciphertext, PublicKey, SecretKey, ThresholdPublicKey, ThresholdSecretKey,
    PartialDecryption, PartialDecryptionZKP, from_records, to_records,
)
