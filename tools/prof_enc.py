"""Small EncryptWithR + Decrypt run for ncu captures: python tools/prof_enc.py [count]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paillier_b200 import synth
from paillier_b200.api import SecretKey
count = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
p, q = synth.load_key("paillier_2048")
sk = SecretKey(p * q, p=p, q=q)
m = synth.plaintexts(count, sk.N, sk.w_n)
r = synth.randomness(count, sk.N, sk.w_n)
c = sk.encrypt_with_r_records(m, r)
d = sk.decrypt_records(c)
assert np.array_equal(d, m)
print("ok", count)
