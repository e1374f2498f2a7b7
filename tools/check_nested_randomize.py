"""NestedRandomize on a SecretKey context (exponentiations over the prime powers) against Python ints at several batch sizes"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paillier_b200 import synth
from paillier_b200.api import Ciphertext, ENC_LEVEL_TWO, SecretKey, PublicKey
from math import gcd

for name in ("paillier_64", "paillier_1024"):
    p, q = synth.load_key(name)
    n = p * q; n2 = n * n; n3 = n2 * n
    sk = SecretKey(n, p=p, q=q)
    rnd = random.Random(5)
    unit = lambda: next(r for r in iter(lambda: rnd.randrange(1, n), None) if gcd(r, n) == 1)
    for count in (1, 3, 8, 9, 30, 33, 100):
        cts = [Ciphertext(rnd.randrange(1, n3), ENC_LEVEL_TWO) for _ in range(count)]
        As = [unit() for _ in range(count)]; Bs = [unit() for _ in range(count)]
        got = [c.C for c in sk.NestedRandomizeWithBatch(cts, As, Bs)]
        want = [pow(c.C, pow(a, n, n2), n3) * pow(b, n2, n3) % n3 for c, a, b in zip(cts, As, Bs)]
        bad = [i for i in range(count) if got[i] != want[i]]
        print(name, "nested_randomize count", count, "bad", bad[:10], len(bad))
    sk.close()
