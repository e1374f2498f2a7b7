"""level-2 key-holder encryption / CRT decryption rates (moduli p^3, q^3: the 96-limb kernel shape at 2048-bit n):
PGPU_SHAPE_96=4,24 python tools/shape96.py [count]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paillier_b200 import synth
from paillier_b200._lib import check, lib
from paillier_b200.api import SecretKey
count = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
p, q = synth.load_key("paillier_2048")
n = p * q
sk = SecretKey(n, p=p, q=q)
m = synth.random_records(count, sk.w_n2, (n * n).bit_length() - 1, stream=51)
r = synth.randomness(count, n, sk.w_n)
c = np.empty(count * sk.w_n3, dtype=np.uint8); d = np.empty(count * sk.w_n2, dtype=np.uint8)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
for timed in (0, 1):
    t0 = time.perf_counter()
    check(lib.pgpu_encrypt_with_r_at_level_sk(sk._ctx, 2, count, vp(m), vp(r), vp(c)), sk._ctx)
    t1 = time.perf_counter()
    check(lib.pgpu_decrypt_at_level(sk._ctx, 2, count, vp(c), vp(d)), sk._ctx)
    t2 = time.perf_counter()
assert np.array_equal(d, m)
print(f"PGPU_SHAPE_96={os.environ.get('PGPU_SHAPE_96', 'default (4,24; 8,12 for per-item exponents)')}: level-2 sk enc {count/(t1-t0):.0f}/s, CRT dec {count/(t2-t1):.0f}/s")
