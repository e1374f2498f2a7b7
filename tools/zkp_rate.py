"""PartialDecrypt / PartialDecryptionWithZKP / VerifyProof rates at a given key size: python tools/zkp_rate.py [count] [bits]"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paillier_b200 import synth
from paillier_b200.keygen import ThresholdKeyGenerator
count = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 3072
p, q = synth.load_key(f"threshold_{bits}")
keys = ThresholdKeyGenerator(bits, 8, 5, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys()
tsk = keys[0]
n = p * q
c = tsk.encrypt_with_r_records(synth.plaintexts(count, n, tsk.w_n), synth.randomness(count, n, tsk.w_n))
r = synth.random_records(count, tsk.w_n2, (n * n).bit_length() - 1, stream=5)
tsk.zkp_prove_records(c[:64 * tsk.w_n2], r[:64 * tsk.w_n2])
l0 = tsk.launch_count()
t0 = time.perf_counter(); tsk.partial_decrypt_records(c); t1 = time.perf_counter()
dec, e, z = tsk.zkp_prove_records(c, r); t2 = time.perf_counter()
l1 = tsk.launch_count()
ok = tsk.verify_proof_records(tsk.ID, c, dec, e, z); t3 = time.perf_counter()
assert ok.all()
print(f"bits {bits} count {count}: pdec {count/(t1-t0):.0f}/s  prove {count/(t2-t1):.0f}/s  verify {count/(t3-t2):.0f}/s  "
      f"prove+verify {count/(t3-t1):.0f}/s  launches prove {l1-l0} verify {tsk.launch_count()-l1}")
