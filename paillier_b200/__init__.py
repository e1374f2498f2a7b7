"""paillier_b200: B200-native batch engine behind sachaservan/paillier's mod-n^2 exponentiation path.

The product is the C-ABI shared library (include/pgpu.h, paillier_b200/csrc); this package is the
host-side mirror of the reference's interface used by the tests and the benchmark.

The API names are resolved lazily (PEP 562): `paillier_b200.synth` (seeded inputs, numpy only) can be imported by
bench.py's CPU reference arm without mapping libpaillier_b200.so into that process; anything that computes
(`paillier_b200.api`, `._lib`) loads the CUDA library and fails loudly if it is missing.
"""
_API = ("Ciphertext", "PublicKey", "SecretKey", "ThresholdPublicKey", "ThresholdSecretKey",
        "PartialDecryption", "PartialDecryptionZKP", "from_records", "to_records")


def __getattr__(name):
    if name in _API:
        from . import api
        return getattr(api, name)
    raise AttributeError(f"module 'paillier_b200' has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + list(_API))
