"""ctypes front-end of oracle/ossl_ref.c (OpenSSL BN): a third bignum implementation for the large-size vectors.

TEST INFRASTRUCTURE ONLY -- see the header of ossl_ref.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libosslref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "ossl_ref.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_SO)
    return _lib


def _be(x: int) -> bytes:
    return x.to_bytes((x.bit_length() + 7) // 8, "big")


def _rec(x: int, w: int) -> np.ndarray:
    return np.frombuffer(x.to_bytes(w, "little"), dtype=np.uint8).copy()


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def exp(x: int, y: int, m):
    """ncw/gmp Int.Exp semantics (y <= 0 -> 1; nil modulus -> plain power) with BN_mod_exp underneath"""
    if y <= 0:
        return 1
    if m is None or m == 0:
        return x ** y
    m = abs(m)
    w = (m.bit_length() + 7) // 8
    eb = max(1, (y.bit_length() + 7) // 8)
    out = np.zeros(w, dtype=np.uint8)
    mb = _be(m)
    rc = lib().ossl_modexp(mb, C.c_size_t(len(mb)), C.c_size_t(1), _p(_rec(x % m, w)), C.c_size_t(w), _p(_rec(y, eb)), C.c_size_t(eb), _p(out))
    if rc:
        raise RuntimeError(f"ossl_modexp failed ({rc})")
    return int.from_bytes(out.tobytes(), "little")


def mod_inverse(a: int, m: int) -> int:
    w = (m.bit_length() + 7) // 8
    out = np.zeros(w, dtype=np.uint8)
    ok = np.zeros(1, dtype=np.uint8)
    mb = _be(m)
    lib().ossl_modinv(mb, C.c_size_t(len(mb)), C.c_size_t(1), _p(_rec(a % m, w)), C.c_size_t(w), _p(out), _p(ok))
    if not ok[0]:
        raise ValueError("base is not invertible for the given modulus")
    return int.from_bytes(out.tobytes(), "little")


def modexp_records(mod: int, base: np.ndarray, width: int, exp_: np.ndarray, exp_bytes: int) -> np.ndarray:
    base = np.ascontiguousarray(base).view(np.uint8).reshape(-1)
    exp_ = np.ascontiguousarray(exp_).view(np.uint8).reshape(-1)
    count = base.size // width
    out = np.zeros(count * width, dtype=np.uint8)
    mb = _be(mod)
    rc = lib().ossl_modexp(mb, C.c_size_t(len(mb)), C.c_size_t(count), _p(base), C.c_size_t(width), _p(exp_), C.c_size_t(exp_bytes), _p(out))
    if rc:
        raise RuntimeError(f"ossl_modexp failed ({rc})")
    return out
