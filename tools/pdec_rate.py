"""PartialDecrypt rate at a given key size: python tools/pdec_rate.py [bits] [count]   (PGPU_SHAPE_<S>=tpi,L picks the kernel shape)"""
import ctypes as C, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paillier_b200 import synth
from paillier_b200._lib import check, lib
from paillier_b200.keygen import ThresholdKeyGenerator
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
count = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
p, q = synth.load_key(f"threshold_{bits}")
n = p * q
keys = ThresholdKeyGenerator(bits, 8, 5, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys()
tsk = keys[0]
dev = torch.device("cuda", 0)
c = torch.from_numpy(synth.random_records(count, tsk.w_n2, (n * n).bit_length() - 1, stream=3)).to(dev)
out = torch.empty_like(c)
vp = lambda t: C.c_void_p(t.data_ptr())
for i in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    check(lib.pgpu_partial_decrypt_dev(tsk._ctx, count, vp(c), vp(out)), tsk._ctx)
    check(lib.pgpu_ctx_set_stream(tsk._ctx, None), tsk._ctx)
    lib.pgpu_ctx_destroy  # noqa
    # the context's private stream: wait for it through a blocking host call
    h = np.empty(tsk.w_n2, dtype=np.uint8)
    check(lib.pgpu_partial_decrypt(tsk._ctx, 1, C.c_void_p(c[:tsk.w_n2].cpu().numpy().ctypes.data), h.ctypes.data_as(C.c_void_p)), tsk._ctx)
    dt = time.perf_counter() - t0
S, sq, mu = tsk.program_cost(2)
print(f"bits {bits} count {count} shape_env {os.environ.get('PGPU_SHAPE_%d' % S)} S {S} sqr {sq} mul {mu}: {count / dt:.0f} pdec/s, "
      f"{count / dt * (sq * (1.5 * S * S + 1.5 * S) + mu * (2 * S * S + S)) / 1e12:.2f} T MAC32/s algorithmic")
