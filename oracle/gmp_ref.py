"""ctypes front-end of oracle/gmp_ref.c (libgmp restatement of the reference's call sequences).

TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see the header of gmp_ref.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgmpref.so")


def build() -> str:
    src = os.path.join(_HERE, "gmp_ref.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _be(x: int) -> bytes:
    return x.to_bytes((x.bit_length() + 7) // 8, "big")


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def encrypt_with_r(n: int, m, r, w_n: int, threads: int = 0) -> np.ndarray:
    m, r = _u8(m), _u8(r)
    count = m.size // w_n
    out = np.zeros(count * 2 * w_n, dtype=np.uint8)
    nb = _be(n)
    lib().ref_encrypt_with_r(nb, C.c_size_t(len(nb)), C.c_size_t(count), _p(m), _p(r), C.c_size_t(w_n), _p(out),
                             C.c_size_t(2 * w_n), threads or cores())
    return out


def decrypt(n: int, lam: int, c, w_n: int, threads: int = 0) -> np.ndarray:
    c = _u8(c)
    count = c.size // (2 * w_n)
    out = np.zeros(count * w_n, dtype=np.uint8)
    nb, lb = _be(n), _be(lam)
    lib().ref_decrypt(nb, C.c_size_t(len(nb)), lb, C.c_size_t(len(lb)), C.c_size_t(count), _p(c), C.c_size_t(2 * w_n),
                      _p(out), C.c_size_t(w_n), threads or cores())
    return out


def partial_decrypt(n: int, share: int, l: int, c, w_n2: int, threads: int = 0) -> np.ndarray:
    c = _u8(c)
    count = c.size // w_n2
    out = np.zeros(count * w_n2, dtype=np.uint8)
    nb, sb = _be(n), _be(share)
    lib().ref_partial_decrypt(nb, C.c_size_t(len(nb)), sb, C.c_size_t(len(sb)), l, C.c_size_t(count), _p(c), _p(out),
                              C.c_size_t(w_n2), threads or cores())
    return out


def modexp(mod: int, base, width: int, exp, exp_bytes: int, threads: int = 0) -> np.ndarray:
    base, exp = _u8(base), _u8(exp)
    count = base.size // width
    out = np.zeros(count * width, dtype=np.uint8)
    mb = _be(mod)
    lib().ref_modexp(mb, C.c_size_t(len(mb)), C.c_size_t(count), _p(base), C.c_size_t(width), _p(exp),
                     C.c_size_t(exp_bytes), _p(out), threads or cores())
    return out


def modmul(mod: int, a, b, width: int, threads: int = 0) -> np.ndarray:
    a, b = _u8(a), _u8(b)
    count = a.size // width
    out = np.zeros(count * width, dtype=np.uint8)
    mb = _be(mod)
    lib().ref_modmul(mb, C.c_size_t(len(mb)), C.c_size_t(count), _p(a), _p(b), C.c_size_t(width), _p(out), threads or cores())
    return out


def modinv(mod: int, a, width: int, threads: int = 0):
    """-> (inverses, ok): mpz_invert per record; ok[i] = 0 where a[i] is not a unit (record left zero)"""
    a = _u8(a)
    count = a.size // width
    out = np.zeros(count * width, dtype=np.uint8)
    ok = np.zeros(count, dtype=np.uint8)
    mb = _be(mod)
    lib().ref_modinv(mb, C.c_size_t(len(mb)), C.c_size_t(count), _p(a), C.c_size_t(width), _p(out), _p(ok), threads or cores())
    return out, ok


def add_reduce(mod: int, c, width: int, threads: int = 0) -> np.ndarray:
    c = _u8(c)
    count = c.size // width
    out = np.zeros(width, dtype=np.uint8)
    mb = _be(mod)
    lib().ref_add_reduce(mb, C.c_size_t(len(mb)), C.c_size_t(count), _p(c), C.c_size_t(width), _p(out), threads or cores())
    return out


def pdec_zkp(n: int, share: int, l: int, v: int, c, r, w_n2: int, w_z: int, threads: int = 0):
    """PartialDecryptionWithZKP (thresholdkey.go:225-255) with r supplied -> (dec, e, z) records (e: 32 bytes little-endian)"""
    c, r = _u8(c), _u8(r)
    count = c.size // w_n2
    dec = np.zeros(count * w_n2, dtype=np.uint8)
    e = np.zeros(count * 32, dtype=np.uint8)
    z = np.zeros(count * w_z, dtype=np.uint8)
    nb, sb, vb = _be(n), _be(share), _be(v)
    lib().ref_pdec_zkp(nb, C.c_size_t(len(nb)), sb, C.c_size_t(len(sb)), l, vb, C.c_size_t(len(vb)), C.c_size_t(count), _p(c), _p(r),
                       C.c_size_t(w_n2), _p(dec), _p(e), _p(z), C.c_size_t(w_z), threads or cores())
    return dec, e, z


def zkp_verify(n: int, v: int, vi: int, c, dec, e, z, w_n2: int, w_z: int, threads: int = 0) -> np.ndarray:
    """PartialDecryptionZKP.VerifyProof (thresholdkey.go:278-311) for proofs of the server whose verification key is vi"""
    c, dec, e, z = _u8(c), _u8(dec), _u8(e), _u8(z)
    count = c.size // w_n2
    ok = np.zeros(count, dtype=np.uint8)
    nb, vb, vib = _be(n), _be(v), _be(vi)
    lib().ref_zkp_verify(nb, C.c_size_t(len(nb)), vb, C.c_size_t(len(vb)), vib, C.c_size_t(len(vib)), C.c_size_t(count), _p(c), _p(dec),
                         _p(e), _p(z), C.c_size_t(w_n2), C.c_size_t(w_z), _p(ok), threads or cores())
    return ok


def ddleq_verify(n: int, secpar: int, ct1, ct2, x, y, alpha, e, f, w_n: int, w_n2: int, w_n3: int, threads: int = 0) -> np.ndarray:
    """verifyDDLEQProofInstance (ddleq.go:129-153) for count statements x secpar instances -> ok per instance"""
    ct1, ct2, x, y, alpha, e, f = (_u8(a) for a in (ct1, ct2, x, y, alpha, e, f))
    count = ct1.size // w_n3
    ok = np.zeros(count * secpar, dtype=np.uint8)
    nb = _be(n)
    lib().ref_ddleq_verify(nb, C.c_size_t(len(nb)), C.c_size_t(count), C.c_uint(secpar), _p(ct1), _p(ct2), _p(x), _p(y), _p(alpha), _p(e),
                           _p(f), C.c_size_t(w_n), C.c_size_t(w_n2), C.c_size_t(w_n3), _p(ok), threads or cores())
    return ok


def dot_u64(mod: int, c, width: int, k, threads: int = 0) -> np.ndarray:
    """prod c[i]^k[i] mod `mod`: ConstMult (operations.go:58-64) per term folded by Add (operations.go:11-29)"""
    c = _u8(c)
    k = np.ascontiguousarray(k, dtype=np.uint64)
    count = c.size // width
    out = np.zeros(width, dtype=np.uint8)
    mb = _be(mod)
    lib().ref_dot_u64(mb, C.c_size_t(len(mb)), C.c_size_t(count), _p(c), C.c_size_t(width), _p(k), _p(out), threads or cores())
    return out


def safe_prime_scan(p_bits: int, raw: bytes, threads: int = 0) -> np.ndarray:
    """one iteration of runGenPrimeRoutine's loop (safe_prime.go:170-263) per byte string -> accept flags"""
    nb = (p_bits - 1 + 7) // 8
    a = np.frombuffer(raw, dtype=np.uint8)
    count = a.size // nb
    ok = np.zeros(count, dtype=np.uint8)
    lib().ref_safe_prime_scan(C.c_uint(p_bits), C.c_size_t(count), _p(a), _p(ok), threads or cores())
    return ok
