"""Host-side mirror of the reference's interface for the batch path.

The Go package keeps its scalar API; the batch methods a maintainer adds next
to it are mirrored here one-to-one, with the same names and
argument meaning, on top of the same C-ABI symbols the cgo layer binds:

    PublicKey.EncryptWithRBatch      <- PublicKey.EncryptWithR      paillier.go:185-187,206-218
    SecretKey.DecryptBatch           <- SecretKey.Decrypt           paillier.go:292-303
    PublicKey.ConstMultBatch         <- PublicKey.ConstMult         operations.go:58-64
    PublicKey.AddBatch / AddPairs    <- PublicKey.Add               operations.go:11-29
    ThresholdSecretKey.PartialDecryptBatch <- PartialDecrypt        thresholdkey.go:192-201

All arithmetic on batch items happens in libpaillier_b200.so on the GPU; this
module only marshals integers to fixed-width little-endian records.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import MOD_N, MOD_N2, MOD_N3, PgpuError, check, lib  # noqa: F401

ENC_LEVEL_ONE, ENC_LEVEL_TWO = 0, 1          # paillier.go:17-23
REGULAR, ALTERNATIVE, MIXED = 0, 1, 2        # paillier.go:29-39


def to_records(values: Sequence[int], width: int) -> np.ndarray:
    """ints -> contiguous uint8 array of fixed-width little-endian records."""
    buf = bytearray(len(values) * width)
    for i, v in enumerate(values):
        buf[i * width:(i + 1) * width] = int(v).to_bytes(width, "little")
    return np.frombuffer(bytes(buf), dtype=np.uint8).copy() if values else np.zeros(0, dtype=np.uint8)


def from_records(buf, width: int) -> List[int]:
    b = bytes(memoryview(np.ascontiguousarray(buf)).cast("B"))
    return [int.from_bytes(b[i:i + width], "little") for i in range(0, len(b), width)]


def _be(x: int) -> bytes:
    """gmp.Int.Bytes(): minimal big-endian magnitude"""
    return x.to_bytes((x.bit_length() + 7) // 8, "big")


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _as_u8(a, nbytes: int, what: str) -> np.ndarray:
    arr = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    if arr.size != nbytes:
        raise ValueError(f"{what}: expected {nbytes} bytes, got {arr.size}")
    return arr


@dataclass
class Ciphertext:
    """paillier.go:65-69"""
    C: int
    Level: int = ENC_LEVEL_ONE
    EncMethod: int = REGULAR


@dataclass
class PartialDecryption:
    """thresholdkey.go:45-48"""
    ID: int
    Decryption: int


@dataclass
class PartialDecryptionZKP:
    """thresholdkey.go:52-58 (Key is the ThresholdPublicKey the verifier already holds)"""
    ID: int
    Decryption: int
    E: int
    Z: int
    C: int


class PublicKey:
    """paillier.go:46-56 with g = n+1 (paillier.go:147); owns one engine context on `device`."""

    def __init__(self, N: int, device: int = 0):
        self.N = int(N)
        self.G = self.N + 1
        self.device = device
        self._ctx = C.c_void_p()
        nb = _be(self.N)
        check(lib.pgpu_ctx_create(C.byref(self._ctx), device, nb, len(nb)))
        wn, w2, w3 = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(lib.pgpu_ctx_widths(self._ctx, C.byref(wn), C.byref(w2), C.byref(w3)), self._ctx)
        self.w_n, self.w_n2, self.w_n3 = wn.value, w2.value, w3.value

    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            lib.pgpu_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- cached moduli (paillier.go:72-90)
    def GetN2(self) -> int:
        return self.N * self.N

    def GetN3(self) -> int:
        return self.N ** 3

    # -- records API (bytes in, bytes out) --------------------------------
    def encrypt_with_r_records(self, m, r) -> np.ndarray:
        m = np.ascontiguousarray(m).view(np.uint8).reshape(-1)
        count = m.size // self.w_n
        m = _as_u8(m, count * self.w_n, "m")
        r = _as_u8(r, count * self.w_n, "r")
        out = np.empty(count * self.w_n2, dtype=np.uint8)
        check(lib.pgpu_encrypt_with_r(self._ctx, count, _ptr(m), _ptr(r), _ptr(out)), self._ctx)
        return out

    def const_mult_records(self, c, k, k_bytes: int) -> np.ndarray:
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // self.w_n2
        c = _as_u8(c, count * self.w_n2, "c")
        k = _as_u8(k, count * k_bytes, "k")
        out = np.empty(count * self.w_n2, dtype=np.uint8)
        check(lib.pgpu_const_mult(self._ctx, count, _ptr(c), _ptr(k), k_bytes, _ptr(out)), self._ctx)
        return out

    def add_reduce_records(self, c) -> np.ndarray:
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // self.w_n2
        out = np.empty(self.w_n2, dtype=np.uint8)
        check(lib.pgpu_add_reduce(self._ctx, count, _ptr(c) if count else None, _ptr(out)), self._ctx)
        return out

    def add_pairs_records(self, a, b) -> np.ndarray:
        a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        count = a.size // self.w_n2
        b = _as_u8(b, count * self.w_n2, "b")
        out = np.empty(count * self.w_n2, dtype=np.uint8)
        check(lib.pgpu_add_pairs(self._ctx, count, _ptr(a), _ptr(b), _ptr(out)), self._ctx)
        return out

    def dot_u64_records(self, c, k: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // self.w_n2
        k = np.ascontiguousarray(k, dtype=np.uint64)
        if k.size != count:
            raise ValueError("one 64-bit scalar per ciphertext")
        out = np.empty(self.w_n2, dtype=np.uint8)
        check(lib.pgpu_dot_u64(self._ctx, count, _ptr(c) if count else None, _ptr(k) if count else None, _ptr(out)), self._ctx)
        return out

    def modexp_records(self, modsel: int, base, exp, exp_bytes: int, width: int) -> np.ndarray:
        base = np.ascontiguousarray(base).view(np.uint8).reshape(-1)
        count = base.size // width
        exp = _as_u8(exp, count * exp_bytes, "exp")
        out = np.empty(count * width, dtype=np.uint8)
        check(lib.pgpu_modexp(self._ctx, modsel, count, _ptr(base), _ptr(exp), exp_bytes, _ptr(out)), self._ctx)
        return out

    def modexp_shared_records(self, modsel: int, base, exponent: int, width: int) -> np.ndarray:
        base = np.ascontiguousarray(base).view(np.uint8).reshape(-1)
        count = base.size // width
        out = np.empty(count * width, dtype=np.uint8)
        eb = _be(exponent)
        check(lib.pgpu_modexp_shared(self._ctx, modsel, count, _ptr(base), eb, len(eb), _ptr(out)), self._ctx)
        return out

    def modmul_records(self, modsel: int, a, b, width: int) -> np.ndarray:
        a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        count = a.size // width
        b = _as_u8(b, count * width, "b")
        out = np.empty(count * width, dtype=np.uint8)
        check(lib.pgpu_modmul(self._ctx, modsel, count, _ptr(a), _ptr(b), _ptr(out)), self._ctx)
        return out

    # -- reference-shaped batch methods (ints in, Ciphertexts out) ---------
    def EncryptWithRBatch(self, ms: Sequence[int], rs: Sequence[int]) -> List[Ciphertext]:
        """N x PublicKey.EncryptWithR (paillier.go:185-187)"""
        if len(ms) != len(rs):
            raise ValueError("one r per plaintext")
        out = self.encrypt_with_r_records(to_records(ms, self.w_n), to_records(rs, self.w_n))
        return [Ciphertext(c, ENC_LEVEL_ONE, REGULAR) for c in from_records(out, self.w_n2)]

    def ConstMultBatch(self, cts: Sequence[Ciphertext], ks: Sequence[int]) -> List[Ciphertext]:
        """N x PublicKey.ConstMult (operations.go:58-64); k <= 0 gives 1 like gmp's Exp"""
        if len(cts) != len(ks):
            raise ValueError("one scalar per ciphertext")
        kmax = max([int(k) for k in ks] + [1])
        k_bytes = max(4, 4 * ((kmax.bit_length() + 31) // 32))
        out = self.const_mult_records(to_records([c.C for c in cts], self.w_n2),
                                      to_records([max(int(k), 0) for k in ks], k_bytes), k_bytes)
        return [Ciphertext(v, ct.Level, ct.EncMethod) for v, ct in zip(from_records(out, self.w_n2), cts)]

    def AddBatch(self, cts: Sequence[Ciphertext]) -> Ciphertext:
        """PublicKey.Add(cts...) (operations.go:11-29)"""
        out = self.add_reduce_records(to_records([c.C for c in cts], self.w_n2))
        return Ciphertext(from_records(out, self.w_n2)[0], ENC_LEVEL_ONE, MIXED)

    def AddPairs(self, a: Sequence[Ciphertext], b: Sequence[Ciphertext]) -> List[Ciphertext]:
        """N x PublicKey.Add(a_i, b_i)"""
        out = self.add_pairs_records(to_records([c.C for c in a], self.w_n2), to_records([c.C for c in b], self.w_n2))
        return [Ciphertext(v, ENC_LEVEL_ONE, MIXED) for v in from_records(out, self.w_n2)]

    def SubPairs(self, a: Sequence[Ciphertext], b: Sequence[Ciphertext]) -> List[Ciphertext]:
        """N x PublicKey.Sub(a_i, b_i) (operations.go:32-55)"""
        ar, br = to_records([c.C for c in a], self.w_n2), to_records([c.C for c in b], self.w_n2)
        out = np.empty(len(a) * self.w_n2, dtype=np.uint8)
        check(lib.pgpu_sub_pairs(self._ctx, len(a), _ptr(ar), _ptr(br), _ptr(out)), self._ctx)
        return [Ciphertext(v, ENC_LEVEL_ONE, MIXED) for v in from_records(out, self.w_n2)]

    def ModInverseBatch(self, xs: Sequence[int], modsel: int = MOD_N2) -> List[int]:
        """N x gmp.Int.ModInverse(x, mod); raises PgpuError(PGPU_ERR_NOT_INVERTIBLE) for a non-unit"""
        width = {MOD_N2: self.w_n2, MOD_N3: self.w_n3}[modsel]
        xr = to_records(xs, width)
        out = np.empty(len(xs) * width, dtype=np.uint8)
        check(lib.pgpu_modinv(self._ctx, modsel, len(xs), _ptr(xr), _ptr(out)), self._ctx)
        return from_records(out, width)

    def DotProduct(self, cts: Sequence[Ciphertext], ks: Sequence[int]) -> Ciphertext:
        """Add(ConstMult(c_i, k_i) ...) with 64-bit scalars (BASELINE config 3)"""
        out = self.dot_u64_records(to_records([c.C for c in cts], self.w_n2), np.array([int(k) for k in ks], dtype=np.uint64))
        return Ciphertext(from_records(out, self.w_n2)[0], ENC_LEVEL_ONE, MIXED)

    def ExpBatch(self, bases: Sequence[int], exps: Sequence[int], modsel: int = MOD_N2) -> List[int]:
        """N x gmp.Int.Exp(base, exp, mod) with per-item exponents"""
        width = {MOD_N2: self.w_n2, MOD_N3: self.w_n3}[modsel]
        emax = max([int(e) for e in exps] + [1])
        eb = max(4, 4 * ((emax.bit_length() + 31) // 32))
        out = self.modexp_records(modsel, to_records(bases, width), to_records([max(int(e), 0) for e in exps], eb), eb, width)
        return from_records(out, width)

    def ExpSharedBatch(self, bases: Sequence[int], exponent: int, modsel: int = MOD_N2) -> List[int]:
        width = {MOD_N2: self.w_n2, MOD_N3: self.w_n3}[modsel]
        return from_records(self.modexp_shared_records(modsel, to_records(bases, width), max(int(exponent), 0), width), width)

    def MulModBatch(self, a: Sequence[int], b: Sequence[int], modsel: int = MOD_N2) -> List[int]:
        width = {MOD_N2: self.w_n2, MOD_N3: self.w_n3}[modsel]
        return from_records(self.modmul_records(modsel, to_records(a, width), to_records(b, width), width), width)

    # -- introspection -------------------------------------------------------
    def launch_count(self) -> int:
        v = C.c_uint64()
        check(lib.pgpu_ctx_launch_count(self._ctx, C.byref(v)), self._ctx)
        return v.value

    def program_cost(self, what: int):
        s, q, m = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib.pgpu_ctx_program_cost(self._ctx, what, C.byref(s), C.byref(q), C.byref(m)), self._ctx)
        return s.value, q.value, m.value


class SecretKey(PublicKey):
    """paillier.go:59-62: the reference keeps Lambda = (p-1)(q-1) only; either form is accepted."""

    def __init__(self, N: int, Lambda: Optional[int] = None, p: Optional[int] = None, q: Optional[int] = None, device: int = 0):
        super().__init__(N, device)
        if p is not None and q is not None:
            pb, qb = _be(p), _be(q)
            check(lib.pgpu_ctx_set_secret_pq(self._ctx, pb, len(pb), qb, len(qb)), self._ctx)
            self.Lambda = (p - 1) * (q - 1)
        elif Lambda is not None:
            lb = _be(Lambda)
            check(lib.pgpu_ctx_set_secret_lambda(self._ctx, lb, len(lb)), self._ctx)
            self.Lambda = Lambda
        else:
            raise ValueError("SecretKey needs Lambda or (p, q)")

    def decrypt_records(self, c) -> np.ndarray:
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // self.w_n2
        c = _as_u8(c, count * self.w_n2, "c")
        out = np.empty(count * self.w_n, dtype=np.uint8)
        check(lib.pgpu_decrypt(self._ctx, count, _ptr(c), _ptr(out)), self._ctx)
        return out

    def DecryptBatch(self, cts: Sequence[Ciphertext]) -> List[int]:
        """N x SecretKey.Decrypt (paillier.go:292-303), level 1"""
        if any(c.Level != ENC_LEVEL_ONE for c in cts):
            raise ValueError("DecryptBatch handles level-1 ciphertexts")
        return from_records(self.decrypt_records(to_records([c.C for c in cts], self.w_n2)), self.w_n)


class ThresholdPublicKey(PublicKey):
    """thresholdkey.go:26-32"""

    def __init__(self, N: int, TotalNumberOfDecryptionServers: int, Threshold: int, VerificationKey: int,
                 VerificationKeys: Sequence[int], device: int = 0, _id: int = 0, _share: Optional[int] = None):
        super().__init__(N, device)
        self.TotalNumberOfDecryptionServers = TotalNumberOfDecryptionServers
        self.Threshold = Threshold
        self.VerificationKey = VerificationKey
        self.VerificationKeys = list(VerificationKeys)
        vb = _be(VerificationKey)
        vk = to_records(self.VerificationKeys, self.w_n2) if self.VerificationKeys else None
        sb = _be(_share) if _share is not None else None
        check(lib.pgpu_ctx_set_threshold(self._ctx, TotalNumberOfDecryptionServers, Threshold, _id,
                                         sb, len(sb) if sb is not None else 0, vb, len(vb),
                                         _ptr(vk) if vk is not None else None), self._ctx)
        wz = C.c_size_t()
        check(lib.pgpu_ctx_z_width(self._ctx, C.byref(wz)), self._ctx)
        self.w_z = wz.value

    def verify_proof_records(self, ID: int, c, dec, e, z) -> np.ndarray:
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // self.w_n2
        dec, e, z = _as_u8(dec, count * self.w_n2, "dec"), _as_u8(e, count * 32, "e"), _as_u8(z, count * self.w_z, "z")
        ok = np.zeros(count, dtype=np.uint8)
        check(lib.pgpu_pdec_zkp_verify(self._ctx, count, ID, _ptr(c), _ptr(dec), _ptr(e), _ptr(z), _ptr(ok)), self._ctx)
        return ok

    def VerifyProofBatch(self, proofs: Sequence[PartialDecryptionZKP]) -> List[bool]:
        """N x PartialDecryptionZKP.VerifyProof (thresholdkey.go:278-291); proofs of one server per call"""
        if not proofs:
            return []
        ids = {p.ID for p in proofs}
        if len(ids) != 1:
            raise ValueError("VerifyProofBatch: one server id per batch")
        ok = self.verify_proof_records(proofs[0].ID, to_records([p.C for p in proofs], self.w_n2),
                                       to_records([p.Decryption for p in proofs], self.w_n2),
                                       to_records([p.E for p in proofs], 32), to_records([p.Z for p in proofs], self.w_z))
        return [bool(x) for x in ok]

    def combine_records(self, ids: Sequence[int], decs) -> np.ndarray:
        """decs: len(ids) consecutive batches of n2-width records (share j's batch first to last)"""
        decs = np.ascontiguousarray(decs).view(np.uint8).reshape(-1)
        k = len(ids)
        count = decs.size // (self.w_n2 * k) if k else 0
        out = np.empty(count * self.w_n, dtype=np.uint8)
        idarr = (C.c_int * max(k, 1))(*ids)
        check(lib.pgpu_combine(self._ctx, count, k, idarr, _ptr(decs) if decs.size else None, _ptr(out) if count else None), self._ctx)
        return out

    def CombinePartialDecryptionsBatch(self, shares: Sequence[Sequence[PartialDecryption]]) -> List[int]:
        """N x CombinePartialDecryptions (thresholdkey.go:149-161): shares[j] is server j's batch, all batches in
        the same ciphertext order.  Raises PgpuError(PGPU_ERR_THRESHOLD) like the reference's errors (:77-89)."""
        ids = [s[0].ID if len(s) else 0 for s in shares]
        flat = [pd.Decryption for s in shares for pd in s]
        return from_records(self.combine_records(ids, to_records(flat, self.w_n2)), self.w_n)

    def CombinePartialDecryptionsZKPBatch(self, shares: Sequence[Sequence[PartialDecryptionZKP]]) -> List[int]:
        """thresholdkey.go:164-172 for a batch: a server's batch takes part only if all its proofs verify."""
        good = [s for s in shares if all(self.VerifyProofBatch(s))]
        return self.CombinePartialDecryptionsBatch([[PartialDecryption(p.ID, p.Decryption) for p in s] for s in good])


class ThresholdSecretKey(ThresholdPublicKey):
    """thresholdkey.go:38-42"""

    def __init__(self, N: int, TotalNumberOfDecryptionServers: int, Threshold: int, VerificationKey: int,
                 VerificationKeys: Sequence[int], ID: int, Share: int, device: int = 0):
        super().__init__(N, TotalNumberOfDecryptionServers, Threshold, VerificationKey, VerificationKeys,
                         device, _id=ID, _share=Share)
        self.ID = ID
        self.Share = Share

    def partial_decrypt_records(self, c) -> np.ndarray:
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // self.w_n2
        out = np.empty(count * self.w_n2, dtype=np.uint8)
        check(lib.pgpu_partial_decrypt(self._ctx, count, _ptr(c), _ptr(out)), self._ctx)
        return out

    def zkp_prove_records(self, c, r):
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // self.w_n2
        r = _as_u8(r, count * self.w_n2, "r")
        dec = np.empty(count * self.w_n2, dtype=np.uint8)
        e = np.empty(count * 32, dtype=np.uint8)
        z = np.empty(count * self.w_z, dtype=np.uint8)
        check(lib.pgpu_pdec_zkp_prove(self._ctx, count, _ptr(c), _ptr(r), _ptr(dec), _ptr(e), _ptr(z)), self._ctx)
        return dec, e, z

    def PartialDecryptionWithZKPBatch(self, cs: Sequence[int], rs: Sequence[int]) -> List[PartialDecryptionZKP]:
        """N x PartialDecryptionWithZKP (thresholdkey.go:225-255); rs are the r in [0, n^2) the reference draws at :233"""
        dec, e, z = self.zkp_prove_records(to_records(cs, self.w_n2), to_records(rs, self.w_n2))
        return [PartialDecryptionZKP(self.ID, d, ee, zz, c) for d, ee, zz, c in
                zip(from_records(dec, self.w_n2), from_records(e, 32), from_records(z, self.w_z), cs)]

    def PartialDecryptBatch(self, cs: Sequence[int]) -> List[PartialDecryption]:
        """N x ThresholdSecretKey.PartialDecrypt (thresholdkey.go:192-201)"""
        out = self.partial_decrypt_records(to_records(cs, self.w_n2))
        return [PartialDecryption(self.ID, v) for v in from_records(out, self.w_n2)]
