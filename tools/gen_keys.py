"""Generates tests/golden/keys.json: the key fixtures every parity test and the benchmark use.

Primes come from `openssl prime -generate [-safe]` (fresh randomness; the result is committed, which
is what makes the fixtures reproducible).  Each prime is re-checked with sympy.  Run once:
    python tools/gen_keys.py
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import sympy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gen_prime(bits, safe, mod4_3=False):
    while True:
        out = subprocess.run(["openssl", "prime", "-generate", "-bits", str(bits)] + (["-safe"] if safe else []),
                             capture_output=True, text=True, check=True).stdout.strip()
        p = int(out)
        if p.bit_length() != bits or (p >> (bits - 2)) != 3:   # top two bits set, like rand.Prime / safe_prime.go:183-190
            continue
        if mod4_3 and p % 4 != 3:
            continue
        assert sympy.isprime(p) and (not safe or sympy.isprime((p - 1) // 2))
        return p


def pair(bits, safe):
    with ThreadPoolExecutor(2) as ex:
        a, b = ex.map(lambda _: gen_prime(bits, safe, mod4_3=True), range(2))
    assert a != b
    return {"p": hex(a), "q": hex(b)}


def main():
    keys = {
        # paillier.KeyGen-style keys: p, q = 3 mod 4 (paillier.go:131-137)
        "paillier_64": pair(32, False),
        "paillier_1024": pair(512, False),      # BASELINE config 1
        "paillier_2048": pair(1024, False),     # BASELINE configs 2, 3
        # threshold keys from safe primes (thresholdkey_generator.go:88-99); safe primes are 3 mod 4 too
        "threshold_512": pair(256, True),
        "threshold_2048": pair(1024, True),     # BASELINE config 5 size
        "threshold_3072": pair(1536, True),     # BASELINE config 4
    }
    path = os.path.join(ROOT, "tests", "golden", "keys.json")
    with open(path, "w") as f:
        json.dump(keys, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
