"""DDLEQ prove / verify rates: [PGPU_SHAPE_96=8,12] python tools/ddleq_rate.py [statements] [secpar]"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from math import gcd
from paillier_b200 import synth
from paillier_b200.api import ENC_LEVEL_TWO, SecretKey
dn = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
secpar = int(sys.argv[2]) if len(sys.argv) > 2 else 8
p, q = synth.load_key("paillier_2048")
n = p * q
sk = SecretKey(n, p=p, q=q)
rnd = random.Random(3)
def units(k):
    out = []
    while len(out) < k:
        r = rnd.randrange(1, n)
        if gcd(r, n) == 1:
            out.append(r)
    return out
inner = sk.EncryptWithRBatch([rnd.randrange(n) for _ in range(dn)], units(dn))
ct1 = sk.EncryptWithRAtLevelBatch([c.C for c in inner], units(dn), ENC_LEVEL_TWO)
As, Bs = units(dn), units(dn)
ct2 = sk.NestedRandomizeWithBatch(ct1, As, Bs)
xs = [units(secpar) for _ in range(dn)]; ys = [units(secpar) for _ in range(dn)]
sk.ProveDDLEQBatch(secpar, ct1[:2], ct2[:2], As[:2], Bs[:2], xs[:2], ys[:2])
best_p = best_v = 1e9
for _ in range(2):
    t0 = time.perf_counter(); proofs = sk.ProveDDLEQBatch(secpar, ct1, ct2, As, Bs, xs, ys); t1 = time.perf_counter()
    ok = sk.VerifyDDLEQProofBatch(ct1, ct2, proofs); t2 = time.perf_counter()
    assert all(ok)
    best_p, best_v = min(best_p, t1 - t0), min(best_v, t2 - t1)
print(f"PGPU_SHAPE_96={os.environ.get('PGPU_SHAPE_96', 'default')}: {dn}x{secpar} prove {dn*secpar/best_p:.0f} inst/s, verify {dn*secpar/best_v:.0f} inst/s")
