"""The README's usage example, runnable: python tools/readme_example.py"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paillier_b200.keygen import KeyGen, ThresholdKeyGenerator

rnd = random.Random(1)
sk, pk = KeyGen(2048, rng=rnd)
cts = pk.EncryptBatch([1, 2, 3], rand=rnd)
total = pk.AddBatch(pk.ConstMultBatch(cts, [10, 20, 30]))
assert sk.DecryptBatch([total]) == [140]
assert sk.DecryptBatch([pk.DotProduct(cts, [10, 20, 30])]) == [140]
ct = pk.NewCiphertextFromBytes(cts[0].Bytes())
assert ct == cts[0]

keys = ThresholdKeyGenerator(2048, 8, 5, rng=rnd).GenerateKeys()
tcts = keys[0].EncryptBatch([7, 8, 9], rand=rnd)
n2 = keys[0].N ** 2
zkp_rs = [rnd.randrange(n2) for _ in tcts]
parts = [k.PartialDecryptionWithZKPBatch([c.C for c in tcts], zkp_rs) for k in keys[:5]]
assert all(all(keys[0].VerifyProofBatch(p)) for p in parts)
assert keys[0].CombinePartialDecryptionsZKPBatch(parts) == [7, 8, 9]
print("readme example ok")
