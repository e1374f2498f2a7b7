"""SURVEY.md 8(f) rank 1 through the C ABI alone (ctypes + numpy, no torch): device buffers (pgpu_buf_*), pinned host memory
(pgpu_host_*), a chained Encrypt -> ConstMult -> Add -> Decrypt that never leaves the GPU, and the chunked double-buffered
host path of the blocking entry points (forced to many ragged chunks in a child process)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import gmp_ref as G
from paillier_b200 import synth
from paillier_b200._lib import PgpuError, check, lib
from paillier_b200.api import SecretKey, from_records

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Buf:
    def __init__(self, sk, nbytes):
        self.h = C.c_void_p()
        check(lib.pgpu_buf_alloc(sk._ctx, nbytes, C.byref(self.h)), sk._ctx)
        self.ptr = C.c_void_p(lib.pgpu_buf_ptr(self.h))

    def up(self, a):
        a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        check(lib.pgpu_buf_upload(self.h, 0, a.ctypes.data_as(C.c_void_p), a.size))
        return self

    def down(self, nbytes):
        out = np.empty(nbytes, dtype=np.uint8)
        check(lib.pgpu_buf_download(self.h, 0, out.ctypes.data_as(C.c_void_p), nbytes))
        return out

    def free(self):
        check(lib.pgpu_buf_free(self.h))


def test_chain_on_device_buffers_only():
    p, q = synth.load_key("paillier_2048")
    n = p * q
    sk = SecretKey(n, p=p, q=q)
    count = 1500
    m = synth.plaintexts(count, n, sk.w_n)
    r = synth.randomness(count, n, sk.w_n)
    k = synth.scalars_u64(count)
    bm, br, bk = Buf(sk, count * sk.w_n).up(m), Buf(sk, count * sk.w_n).up(r), Buf(sk, count * 8).up(k)
    bc, bc2, bt, bo = Buf(sk, count * sk.w_n2), Buf(sk, count * sk.w_n2), Buf(sk, sk.w_n2), Buf(sk, sk.w_n)
    assert lib.pgpu_buf_size(bc.h) == count * sk.w_n2
    check(lib.pgpu_encrypt_with_r_dev(sk._ctx, count, bm.ptr, br.ptr, bc.ptr), sk._ctx)
    check(lib.pgpu_const_mult_dev(sk._ctx, count, bc.ptr, bk.ptr, 8, bc2.ptr), sk._ctx)          # k_i * m_i
    check(lib.pgpu_add_pairs_dev(sk._ctx, count, bc2.ptr, bc.ptr, bc2.ptr), sk._ctx)             # + m_i
    check(lib.pgpu_add_reduce_dev(sk._ctx, count, bc2.ptr, bt.ptr), sk._ctx)
    check(lib.pgpu_decrypt_dev(sk._ctx, 1, bt.ptr, bo.ptr), sk._ctx)
    check(lib.pgpu_ctx_sync(sk._ctx), sk._ctx)
    ms = from_records(m, sk.w_n)
    want = sum((int(ki) + 1) * mi for ki, mi in zip(k, ms)) % n
    assert from_records(bo.down(sk.w_n), sk.w_n) == [want]
    # the ciphertexts on the device are the reference's (libgmp call sequence), bit for bit
    assert np.array_equal(bc.down(64 * sk.w_n2), G.encrypt_with_r(n, m[:64 * sk.w_n], r[:64 * sk.w_n], sk.w_n))
    # range checks
    with pytest.raises(PgpuError):
        check(lib.pgpu_buf_upload(bo.h, 8, m.ctypes.data_as(C.c_void_p), sk.w_n))
    with pytest.raises(PgpuError):
        check(lib.pgpu_buf_download(bo.h, 0, m.ctypes.data_as(C.c_void_p), sk.w_n + 1))
    for b in (bm, br, bk, bc, bc2, bt, bo):
        b.free()
    sk.close()


def test_pinned_host_buffers_round_trip():
    p, q = synth.load_key("paillier_2048")
    n = p * q
    sk = SecretKey(n, p=p, q=q)
    count = 777
    def pinned(nbytes):
        ptr = C.c_void_p()
        check(lib.pgpu_host_alloc(nbytes, C.byref(ptr)))
        return ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))
    pm, am = pinned(count * sk.w_n); pr, ar = pinned(count * sk.w_n); pc, ac = pinned(count * sk.w_n2); pd, ad = pinned(count * sk.w_n)
    am[:] = synth.plaintexts(count, n, sk.w_n); ar[:] = synth.randomness(count, n, sk.w_n)
    check(lib.pgpu_encrypt_with_r(sk._ctx, count, pm, pr, pc), sk._ctx)
    check(lib.pgpu_decrypt(sk._ctx, count, pc, pd), sk._ctx)
    assert np.array_equal(ad, am)
    assert np.array_equal(ac[:32 * sk.w_n2], G.encrypt_with_r(n, am[:32 * sk.w_n].copy(), ar[:32 * sk.w_n].copy(), sk.w_n))
    del am, ar, ac, ad
    for ptr in (pm, pr, pc, pd):
        check(lib.pgpu_host_free(ptr))
    sk.close()


CHILD = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from oracle import gmp_ref as G
from paillier_b200 import synth
from paillier_b200.api import PublicKey, SecretKey
p, q = synth.load_key("paillier_2048"); n = p * q
sk = SecretKey(n, p=p, q=q)
count = 2311                                   # 10 chunks of 250 (rounded to whole grids: see chunk_items) and a ragged tail
m = synth.plaintexts(count, n, sk.w_n); r = synth.randomness(count, n, sk.w_n)
c = PublicKey.encrypt_with_r_records(sk, m, r)              # pgpu_encrypt_with_r
assert np.array_equal(c, G.encrypt_with_r(n, m, r, sk.w_n)), "chunked EncryptWithR differs from libgmp"
assert np.array_equal(sk.decrypt_records(c), m), "chunked Decrypt"
assert np.array_equal(sk.encrypt_with_r_records(m, r), c), "chunked EncryptWithR (key holder)"      # pgpu_encrypt_with_r_sk
k = synth.scalars_u64(count)
cm = sk.const_mult_records(c, k.view(np.uint8), 8)
assert np.array_equal(cm, G.modexp(n * n, c, sk.w_n2, k.view(np.uint8), 8)), "chunked ConstMult"
ap = sk.modmul_records(1, c, cm, sk.w_n2)
assert np.array_equal(ap, G.modmul(n * n, c, cm, sk.w_n2)), "chunked AddPairs"
print("ok")
"""


@pytest.mark.parametrize("chunk", ["250", "1000", "0"])
def test_chunked_host_path_many_ragged_chunks(chunk):
    env = dict(os.environ, PGPU_CHUNK_ITEMS=chunk)
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-3000:]
