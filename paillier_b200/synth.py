"""Seeded synthetic workloads: the same bytes are fed to the GPU path, the oracle and
the CPU baseline.  Keys are the committed fixtures of tests/golden/keys.json (tools/gen_keys.py)."""
from __future__ import annotations

import json
import os

import numpy as np

SEED = 20260101
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_key(name: str):
    """-> (p, q) of a fixture key, e.g. 'paillier_2048', 'threshold_3072'"""
    with open(os.path.join(_ROOT, "tests", "golden", "keys.json")) as f:
        k = json.load(f)[name]
    return int(k["p"], 16), int(k["q"], 16)


def random_records(count: int, width: int, bits: int, seed: int = SEED, stream: int = 0, odd: bool = False) -> np.ndarray:
    """count records of `width` bytes (little-endian), each a uniform integer below 2^bits."""
    rng = np.random.Generator(np.random.PCG64([seed, stream]))
    a = rng.integers(0, 256, size=(count, width), dtype=np.uint8)
    full, rem = divmod(bits, 8)
    if full < width:
        a[:, full + (1 if rem else 0):] = 0
        if rem:
            a[:, full] &= (1 << rem) - 1
    if odd and count:
        a[:, 0] |= 1
    return a.reshape(-1)


def plaintexts(count: int, n: int, w_n: int, seed: int = SEED) -> np.ndarray:
    """m uniform in [0, 2^(bitlen(n)-1)), hence < n"""
    return random_records(count, w_n, n.bit_length() - 1, seed, stream=1)


def randomness(count: int, n: int, w_n: int, seed: int = SEED) -> np.ndarray:
    """r odd and below 2^(bitlen(n)-1): in [1, n) and coprime to n except with negligible probability"""
    return random_records(count, w_n, n.bit_length() - 1, seed, stream=2, odd=True)


def scalars_u64(count: int, seed: int = SEED) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64([seed, 3]))
    return rng.integers(1, 2 ** 64, size=count, dtype=np.uint64)
