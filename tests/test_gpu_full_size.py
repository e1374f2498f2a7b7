"""Full BASELINE sizes through size-independent properties (the oracle finishes only samples of these in seconds):
2^20 EncryptWithR -> CRT Decrypt round trip, additive homomorphism over the whole batch, sampled bit-exactness
against libgmp at the head, the tail and across the last partially filled round of resident groups."""
import numpy as np
import pytest

from oracle import gmp_ref as G
from paillier_b200 import synth
from paillier_b200.api import Ciphertext, PublicKey, SecretKey, from_records

pytestmark = pytest.mark.gpu


def test_config2_full_batch_properties():
    p, q = synth.load_key("paillier_2048")
    n = p * q
    sk = SecretKey(n, p=p, q=q)
    count = 1 << 20
    m = synth.plaintexts(count, n, sk.w_n)
    r = synth.randomness(count, n, sk.w_n)
    c = PublicKey.encrypt_with_r_records(sk, m, r)               # public-key path: r^n mod n^2
    assert np.array_equal(sk.encrypt_with_r_records(m, r), c)    # secret-key path (CRT over p^2, q^2): same 2^20 ciphertexts
    for lo in (0, 9472 * 55 - 16, count - 128):                  # head, a round boundary of the persistent grid, tail
        sl = slice(lo * sk.w_n, (lo + 128) * sk.w_n)
        assert np.array_equal(c[lo * sk.w_n2:(lo + 128) * sk.w_n2], G.encrypt_with_r(n, m[sl], r[sl], sk.w_n))
    d = sk.decrypt_records(c)
    assert np.array_equal(d, m)                                   # encode -> decode round trip, all 2^20 items
    total = sk.add_reduce_records(c)                              # Add over the whole batch
    ms = m.reshape(count, sk.w_n)
    # sum of the plaintexts mod n, computed limb-wise in Python ints on 64-bit columns
    cols = ms.view("<u8").astype(object).sum(axis=0)
    s = sum(int(v) << (64 * i) for i, v in enumerate(cols)) % n
    assert sk.DecryptBatch([Ciphertext(from_records(total, sk.w_n2)[0])]) == [s]
    # config 3: dot product with 64-bit scalars == sum k_i m_i mod n
    k = synth.scalars_u64(count)
    dot = sk.dot_u64_records(c, k)
    acc = 0
    kk = k.astype(object)
    for i in range(ms.shape[1] // 8):
        col = ms.view("<u8")[:, i].astype(object)
        acc += int((col * kk).sum()) << (64 * i)
    assert sk.DecryptBatch([Ciphertext(from_records(dot, sk.w_n2)[0])]) == [acc % n]
    sk.close()
