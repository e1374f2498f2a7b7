import sys, time, os
sys.path.insert(0, os.getcwd())
from paillier_b200 import synth
from paillier_b200.keygen import safe_prime_scan
for n in (256, 32768, 32768, 131072):
    raw = synth.random_records(n, 128, 1024, stream=41).tobytes()
    t0 = time.perf_counter(); ps, qs, ok = safe_prime_scan(1024, raw); dt = time.perf_counter() - t0
    print(n, dt, n / dt, sum(ok))
