"""PartialDecryptionWithZKP + VerifyProof over a batch, for ncu launch lists: python tools/prof_zkp.py [count] [bits]"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paillier_b200 import synth
from paillier_b200.keygen import ThresholdKeyGenerator
count = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
p, q = synth.load_key(f"threshold_{bits}")
keys = ThresholdKeyGenerator(bits, 8, 5, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys()
tsk = keys[0]
n = p * q
c = tsk.encrypt_with_r_records(synth.plaintexts(count, n, tsk.w_n), synth.randomness(count, n, tsk.w_n))
r = synth.random_records(count, tsk.w_n2, (n * n).bit_length() - 1, stream=5)
dec, e, z = tsk.zkp_prove_records(c, r)
ok = tsk.verify_proof_records(tsk.ID, c, dec, e, z)
assert ok.all()
print("ok", count)
