"""Time to find 1024-bit safe primes and to generate a 2048-bit threshold key (8 servers, threshold 5) on the GPU."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paillier_b200.keygen import GenerateSafePrime, ThresholdKeyGenerator
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
reader = lambda nb: rnd.randbytes(nb)
GenerateSafePrime(64, reader, batch=4096)          # warm up
ts = []
for i in range(4):
    t0 = time.perf_counter(); p, q = GenerateSafePrime(1024, reader, batch=1 << 16, max_batches=4096); ts.append(time.perf_counter() - t0)
    assert p == 2 * q + 1 and p.bit_length() == 1024
print("1024-bit safe primes, seconds each:", [round(t, 2) for t in ts])
t0 = time.perf_counter()
keys = ThresholdKeyGenerator(2048, 8, 5, rng=rnd, batch=1 << 16).GenerateKeys()
print("2048-bit threshold KeyGen (two safe primes + shares + verification keys): %.2f s" % (time.perf_counter() - t0), keys[0].N.bit_length())
