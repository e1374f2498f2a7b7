import sys, os
sys.path.insert(0, os.getcwd())
from paillier_b200 import synth
from paillier_b200.keygen import safe_prime_scan
raw = synth.random_records(64, 2, 16, stream=41).tobytes()
print(safe_prime_scan(16, raw)[2][:8])
raw = synth.random_records(300, 128, 1024, stream=41).tobytes()
print(sum(safe_prime_scan(1024, raw)[2]))
