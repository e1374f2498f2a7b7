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
from ._lib import MOD_N, MOD_N2, MOD_N3, PGPU_ERR_THRESHOLD, PGPU_OK, PgpuError, check, lib  # noqa: F401

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

    def Bytes(self) -> bytes:
        """Ciphertext.Bytes (paillier.go:392-401): the encoding/gob stream of the struct (gobwire.py)"""
        from .gobwire import encode_ciphertext
        return encode_ciphertext(self.C, self.Level, self.EncMethod)


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


@dataclass
class DDLEQProofInstance:
    """ddleq.go:11-13"""
    X: int
    Y: int
    Alpha: int
    E: int
    F: int


@dataclass
class DDLEQProof:
    """ddleq.go:15-17"""
    Instances: List[DDLEQProofInstance]


class DeviceBuffer:
    """Device memory on a key's device (pgpu_buf_*): what the *Dev methods take, so that ciphertexts stay on the GPU between
    Encrypt / ConstMult / Add / Decrypt calls (the callers of operations.go:11-64).  Made by `key.NewDeviceBuffer(nbytes)`;
    free it (or let it be collected) before the key it came from is closed.  The same type as Go's DeviceBuffer and
    paillier::DeviceBuffer of the C++ mirror."""

    def __init__(self, ctx, nbytes: int):
        self._h = C.c_void_p()
        check(lib.pgpu_buf_alloc(ctx, nbytes, C.byref(self._h)), ctx)

    @property
    def ptr(self) -> C.c_void_p:
        return C.c_void_p(lib.pgpu_buf_ptr(self._h))

    def __len__(self) -> int:
        return lib.pgpu_buf_size(self._h)

    def Upload(self, records, offset: int = 0) -> "DeviceBuffer":
        a = np.ascontiguousarray(records).view(np.uint8).reshape(-1)
        check(lib.pgpu_buf_upload(self._h, offset, _ptr(a), a.size))
        return self

    def Download(self, nbytes: Optional[int] = None, offset: int = 0) -> np.ndarray:
        out = np.empty(len(self) - offset if nbytes is None else nbytes, dtype=np.uint8)
        check(lib.pgpu_buf_download(self._h, offset, _ptr(out), out.size))
        return out

    def Free(self) -> None:
        if self._h is not None and self._h.value:
            check(lib.pgpu_buf_free(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Free()
        except Exception:
            pass


def _need(buf: DeviceBuffer, nbytes: int) -> C.c_void_p:
    if len(buf) < nbytes:
        raise ValueError("device buffer smaller than the batch")
    return buf.ptr


class PublicKey:
    """paillier.go:46-56 with g = n+1 (paillier.go:147); owns one engine context on `device`.
    H and K (alternative encryption, paillier.go:151,158-166) are optional."""

    def __init__(self, N: int, device: int = 0, H: Optional[int] = None, K: Optional[int] = None):
        self.N = int(N)
        self.G = self.N + 1
        self.H, self.K = H, K
        self.device = device
        self._ctx = C.c_void_p()
        nb = _be(self.N)
        check(lib.pgpu_ctx_create(C.byref(self._ctx), device, nb, len(nb)))
        wn, w2, w3 = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(lib.pgpu_ctx_widths(self._ctx, C.byref(wn), C.byref(w2), C.byref(w3)), self._ctx)
        self.w_n, self.w_n2, self.w_n3 = wn.value, w2.value, w3.value
        wm = C.c_size_t()
        check(lib.pgpu_ctx_mod_width(self._ctx, MOD_N, C.byref(wm)), self._ctx)
        self.w_n_rec = wm.value              # record width of the generic mod-n entry points (the n kernel shape)
        if H is not None:
            if K is None or K < 2 or K & (K - 1):
                raise ValueError("K must be a power of two (paillier.go:151)")
            hb = _be(H)
            check(lib.pgpu_ctx_set_alt_generator(self._ctx, hb, len(hb), K.bit_length() - 1), self._ctx)

    def _level_widths(self, level: int):
        """(plaintext width, ciphertext width) of getModuliForLevel (paillier.go:403-414)"""
        if level == ENC_LEVEL_ONE:
            return self.w_n, self.w_n2
        if level == ENC_LEVEL_TWO:
            return self.w_n2, self.w_n3
        raise ValueError("unknown encryption level")

    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            lib.pgpu_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def NewCiphertextFromBytes(self, data: bytes) -> Ciphertext:
        """PublicKey.NewCiphertextFromBytes (paillier.go:374-390); like the reference it does not range-check C"""
        from .gobwire import decode_ciphertext
        return Ciphertext(*decode_ciphertext(data))

    def CiphertextsToBytes(self, cts: Sequence[Ciphertext]) -> List[bytes]:
        return [c.Bytes() for c in cts]

    def NewCiphertextsFromBytes(self, blobs: Sequence[bytes]) -> List[Ciphertext]:
        return [self.NewCiphertextFromBytes(b) for b in blobs]

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

    def precompute_rn_records(self, r) -> np.ndarray:
        """r^n mod n^2 for a pool of r values (the offline half of EncryptWithR): EncryptWithR with m = 0."""
        r = np.ascontiguousarray(r).view(np.uint8).reshape(-1)
        return self.encrypt_with_r_records(np.zeros(r.size, dtype=np.uint8), r)

    def encrypt_with_rn_records(self, m, rn) -> np.ndarray:
        """online half: c = (1 + m*n) * rn mod n^2 (pgpu_encrypt_with_rn)"""
        m = np.ascontiguousarray(m).view(np.uint8).reshape(-1)
        count = m.size // self.w_n
        m = _as_u8(m, count * self.w_n, "m")
        rn = _as_u8(rn, count * self.w_n2, "rn")
        out = np.empty(count * self.w_n2, dtype=np.uint8)
        check(lib.pgpu_encrypt_with_rn(self._ctx, count, _ptr(m), _ptr(rn), _ptr(out)), self._ctx)
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

    def PrecomputeRnBatch(self, rs: Sequence[int]) -> List[int]:
        """r^n mod n^2 for later EncryptWithRnBatch calls (each value is one ciphertext's randomness: use it once)"""
        return from_records(self.precompute_rn_records(to_records(rs, self.w_n)), self.w_n2)

    def EncryptWithRnBatch(self, ms: Sequence[int], rns: Sequence[int]) -> List[Ciphertext]:
        """N x EncryptWithR (paillier.go:206-218) with r^n already computed: two multiplications per item"""
        if len(ms) != len(rns):
            raise ValueError("one r^n per plaintext")
        out = self.encrypt_with_rn_records(to_records(ms, self.w_n), to_records(rns, self.w_n2))
        return [Ciphertext(c, ENC_LEVEL_ONE, REGULAR) for c in from_records(out, self.w_n2)]

    def _draw_units(self, count: int, rand=None) -> List[int]:
        """count x GetRandomNumberInMultiplicativeGroup(n) (utils.go:36-49: uniform below n, redrawn on 0 or gcd(n, r) != 1).
        The draws come from the host CSPRNG (`secrets`, or `rand.randrange` if given); the unit test of the whole batch is
        one batched ModInverse mod n on the GPU, the host's gcd only names the culprits after a failure."""
        import secrets
        draw = (lambda: rand.randrange(self.N)) if rand is not None else (lambda: secrets.randbelow(self.N))
        rs = [draw() for _ in range(count)]
        while count:
            bad = [i for i, r in enumerate(rs) if r == 0]
            if not bad:
                try:
                    self.ModInverseBatch(rs, MOD_N)
                    break
                except PgpuError as e:
                    if e.code != _lib.PGPU_ERR_NOT_INVERTIBLE:
                        raise
                    from math import gcd
                    bad = [i for i, r in enumerate(rs) if gcd(r, self.N) != 1]
            for i in bad:
                rs[i] = draw()
        return rs

    def EncryptAtLevelBatch(self, ms: Sequence[int], level: int, rand=None) -> List[Ciphertext]:
        """N x PublicKey.EncryptAtLevel (paillier.go:258-269): r drawn per item, then EncryptWithRAtLevel"""
        return self.EncryptWithRAtLevelBatch(ms, self._draw_units(len(ms), rand), level)

    def EncryptBatch(self, ms: Sequence[int], rand=None) -> List[Ciphertext]:
        """N x PublicKey.Encrypt (paillier.go:192-194) = EncryptAtLevel(m, DefaultEncryptionLevel)"""
        return self.EncryptAtLevelBatch(ms, ENC_LEVEL_ONE, rand)

    def NestedEncryptBatch(self, ms: Sequence[int], rand=None) -> List[Ciphertext]:
        """N x PublicKey.NestedEncrypt (paillier.go:200-203): a level-2 encryption of the level-1 ciphertext"""
        inner = self.EncryptAtLevelBatch(ms, ENC_LEVEL_ONE, rand)
        return self.EncryptAtLevelBatch([c.C for c in inner], ENC_LEVEL_TWO, rand)

    def AltEncryptAtLevelBatch(self, ms: Sequence[int], level: int, rand=None) -> List[Ciphertext]:
        """N x PublicKey.AltEncryptAtLevel (paillier.go:244-255): r drawn from Z*_n as there, reduced mod K by
        AltEncryptWithRAtLevel"""
        return self.AltEncryptWithRAtLevelBatch(ms, self._draw_units(len(ms), rand), level)

    def EncryptZeroAtLevelBatch(self, count: int, level: int, rand=None) -> List[Ciphertext]:
        """count x PublicKey.EncryptZeroAtLevel (paillier.go:282-284)"""
        return self.EncryptAtLevelBatch([0] * count, level, rand)

    def EncryptOneAtLevelBatch(self, count: int, level: int, rand=None) -> List[Ciphertext]:
        """count x PublicKey.EncryptOneAtLevel (paillier.go:287-289)"""
        return self.EncryptAtLevelBatch([1] * count, level, rand)

    def EncryptZeroBatch(self, count: int, rand=None) -> List[Ciphertext]:
        """count x PublicKey.EncryptZero (paillier.go:272-274)"""
        return self.EncryptZeroAtLevelBatch(count, ENC_LEVEL_ONE, rand)

    def EncryptOneBatch(self, count: int, rand=None) -> List[Ciphertext]:
        """count x PublicKey.EncryptOne (paillier.go:277-279)"""
        return self.EncryptOneAtLevelBatch(count, ENC_LEVEL_ONE, rand)

    # operations.go:11-64 take the modulus from the level of the (first) ciphertext: n^2 at level 1, n^3 at level 2
    def _level_modulus(self, level: int):
        """-> (modulus selector, record width, modulus) of getModuliForLevel(level) (paillier.go:403-414)"""
        if level == ENC_LEVEL_ONE:
            return MOD_N2, self.w_n2, self.N ** 2
        if level == ENC_LEVEL_TWO:
            if not self.w_n3:
                raise PgpuError(_lib.PGPU_ERR_UNSUPPORTED, "n^3 is wider than the built kernel shapes")
            return MOD_N3, self.w_n3, self.N ** 3
        raise ValueError("unsupported encryption level")

    @staticmethod
    def _one_level(cts: Sequence[Ciphertext], what: str) -> int:
        levels = {c.Level for c in cts}
        if len(levels) > 1:
            raise ValueError(f"{what}: one encryption level per batch")
        return levels.pop() if levels else ENC_LEVEL_ONE

    def ConstMultBatch(self, cts: Sequence[Ciphertext], ks: Sequence[int]) -> List[Ciphertext]:
        """N x PublicKey.ConstMult (operations.go:58-64) at the ciphertexts' level; k <= 0 gives 1 like gmp's Exp"""
        if len(cts) != len(ks):
            raise ValueError("one scalar per ciphertext")
        modsel, width, mod = self._level_modulus(self._one_level(cts, "ConstMultBatch"))
        kmax = max([int(k) for k in ks] + [1])
        k_bytes = max(4, 4 * ((kmax.bit_length() + 31) // 32))
        base = to_records([c.C % mod for c in cts], width)
        out = np.empty(len(cts) * width, dtype=np.uint8)
        if len(cts):
            check(lib.pgpu_modexp(self._ctx, modsel, len(cts), _ptr(base), _ptr(to_records([max(int(k), 0) for k in ks], k_bytes)), k_bytes, _ptr(out)),
                  self._ctx)
        return [Ciphertext(v, ct.Level, ct.EncMethod) for v, ct in zip(from_records(out, width), cts)]

    def AddBatch(self, cts: Sequence[Ciphertext]) -> Ciphertext:
        """PublicKey.Add(cts...) (operations.go:11-29): the modulus and the result's level are those of cts[0]"""
        level = cts[0].Level if len(cts) else ENC_LEVEL_ONE
        modsel, width, mod = self._level_modulus(level)
        rec = to_records([c.C % mod for c in cts], width)
        out = np.empty(width, dtype=np.uint8)
        check(lib.pgpu_add_reduce_at_level(self._ctx, level + 1, len(cts), _ptr(rec) if len(cts) else None, _ptr(out)), self._ctx)
        return Ciphertext(from_records(out, width)[0], level, MIXED)

    def AddPairs(self, a: Sequence[Ciphertext], b: Sequence[Ciphertext]) -> List[Ciphertext]:
        """N x PublicKey.Add(a_i, b_i): modulus and level of a_i (one level per batch)"""
        level = self._one_level(a, "AddPairs")
        modsel, width, mod = self._level_modulus(level)
        out = self.modmul_records(modsel, to_records([c.C % mod for c in a], width), to_records([c.C % mod for c in b], width), width)
        return [Ciphertext(v, level, MIXED) for v in from_records(out, width)]

    def SubPairs(self, a: Sequence[Ciphertext], b: Sequence[Ciphertext]) -> List[Ciphertext]:
        """N x PublicKey.Sub(a_i, b_i) (operations.go:32-55): modulus and level of a_i (one level per batch)"""
        level = self._one_level(a, "SubPairs")
        modsel, width, mod = self._level_modulus(level)
        ar, br = to_records([c.C % mod for c in a], width), to_records([c.C % mod for c in b], width)
        out = np.empty(len(a) * width, dtype=np.uint8)
        if level == ENC_LEVEL_ONE:
            check(lib.pgpu_sub_pairs(self._ctx, len(a), _ptr(ar), _ptr(br), _ptr(out)), self._ctx)
        elif len(a):
            inv = np.empty_like(br)
            check(lib.pgpu_modinv(self._ctx, modsel, len(a), _ptr(br), _ptr(inv)), self._ctx)
            check(lib.pgpu_modmul(self._ctx, modsel, len(a), _ptr(ar), _ptr(inv), _ptr(out)), self._ctx)
        return [Ciphertext(v, level, MIXED) for v in from_records(out, width)]

    def ModInverseBatch(self, xs: Sequence[int], modsel: int = MOD_N2) -> List[int]:
        """N x gmp.Int.ModInverse(x, mod); raises PgpuError(PGPU_ERR_NOT_INVERTIBLE) for a non-unit"""
        width = {MOD_N: self.w_n_rec, MOD_N2: self.w_n2, MOD_N3: self.w_n3}[modsel]
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

    # -- levels, alternative encryption, nested operations (paillier.go:206-238, operations.go:67-140) --
    def EncryptWithRAtLevelBatch(self, ms: Sequence[int], rs: Sequence[int], level: int) -> List[Ciphertext]:
        """N x PublicKey.EncryptWithRAtLevel (paillier.go:206-218)"""
        wm, wc = self._level_widths(level)
        mr, rr = to_records(ms, wm), to_records(rs, self.w_n)
        out = np.empty(len(ms) * wc, dtype=np.uint8)
        check(lib.pgpu_encrypt_with_r_at_level(self._ctx, level + 1, len(ms), _ptr(mr), _ptr(rr), _ptr(out)), self._ctx)
        return [Ciphertext(c, level, REGULAR) for c in from_records(out, wc)]

    def AltEncryptWithRAtLevelBatch(self, ms: Sequence[int], rs: List[int], level: int) -> List[Ciphertext]:
        """N x PublicKey.AltEncryptWithRAtLevel (paillier.go:221-238).  Like the reference (:228) the caller's
        r values are reduced mod K in place."""
        if self.H is None:
            raise ValueError("AltEncrypt needs PublicKey.H and K")
        wm, wc = self._level_widths(level)
        mr, rr = to_records(ms, wm), to_records([r % (1 << (8 * self.w_n)) for r in rs], self.w_n)   # the engine uses r mod K
        out = np.empty(len(ms) * wc, dtype=np.uint8)
        check(lib.pgpu_alt_encrypt_with_r_at_level(self._ctx, level + 1, len(ms), _ptr(mr), _ptr(rr), _ptr(out)), self._ctx)
        for i in range(len(rs)):
            rs[i] = rs[i] % self.K
        return [Ciphertext(c, level, ALTERNATIVE) for c in from_records(out, wc)]

    def RandomizeWithRBatch(self, cts: Sequence[Ciphertext], rs: Sequence[int]) -> List[Ciphertext]:
        """N x PublicKey.Randomize (operations.go:67-69) = Add(ct, EncryptWithR(0, r)) with r supplied.  Add takes the modulus
        from ct.Level (operations.go:13-15) while the fresh Encrypt(0) is always a level-1 ciphertext: a level-2 ct is
        multiplied by r^n mod n^2 modulo n^3, as the reference does (bit-exact with it; that product is not an encryption of
        the same plaintext -- the reference's Randomize is only meaningful at level 1, NestedRandomize is the level-2 form)."""
        if len(cts) != len(rs):
            raise ValueError("one r per ciphertext")
        level = self._one_level(cts, "RandomizeWithRBatch")
        if level == ENC_LEVEL_TWO:
            zeros = self.EncryptWithRBatch([0] * len(cts), rs)
            modsel, width, mod = self._level_modulus(level)
            vals = self.MulModBatch([c.C % mod for c in cts], [z.C for z in zeros], modsel)
            return [Ciphertext(v, level, MIXED) for v in vals]
        cr, rr = to_records([c.C for c in cts], self.w_n2), to_records(rs, self.w_n)
        out = np.empty(len(cts) * self.w_n2, dtype=np.uint8)
        check(lib.pgpu_randomize_with_r(self._ctx, len(cts), _ptr(cr), _ptr(rr), _ptr(out)), self._ctx)
        return [Ciphertext(c, ENC_LEVEL_ONE, MIXED) for c in from_records(out, self.w_n2)]

    def RandomizeBatch(self, cts: Sequence[Ciphertext], rand=None) -> List[Ciphertext]:
        """N x PublicKey.Randomize (operations.go:67-69) with the r of each fresh Encrypt(0) drawn here"""
        return self.RandomizeWithRBatch(cts, self._draw_units(len(cts), rand))

    def NestedRandomizeBatch(self, cts: Sequence[Ciphertext], rand=None):
        """N x PublicKey.NestedRandomize (operations.go:96-118) -> (randomized ciphertexts, a values, b values) with a, b drawn
        from Z*_n as there (:105-106)"""
        As, Bs = self._draw_units(len(cts), rand), self._draw_units(len(cts), rand)
        return self.NestedRandomizeWithBatch(cts, As, Bs), As, Bs

    def SubBatch(self, cts: Sequence[Ciphertext]) -> Ciphertext:
        """PublicKey.Sub(cts...) (operations.go:32-55): cts[0] * prod_{i>0} cts[i]^-1 modulo n^(s+1) of cts[0].Level.  The
        inverses of the reference's loop are taken once, of the product of cts[1:] (the same canonical residue); like there,
        a single argument comes back unreduced.  Raises PgpuError(PGPU_ERR_NOT_INVERTIBLE) where ModInverse has no result."""
        if not cts:
            raise ValueError("Sub needs at least one ciphertext")                  # the reference indexes cts[0]
        level = cts[0].Level
        if len(cts) == 1:
            return Ciphertext(cts[0].C, level, MIXED)
        modsel, width, mod = self._level_modulus(level)
        rest = to_records([c.C % mod for c in cts[1:]], width)
        prod = np.empty(width, dtype=np.uint8)
        check(lib.pgpu_add_reduce_at_level(self._ctx, level + 1, len(cts) - 1, _ptr(rest), _ptr(prod)), self._ctx)
        inv = np.empty(width, dtype=np.uint8)
        check(lib.pgpu_modinv(self._ctx, modsel, 1, _ptr(prod), _ptr(inv)), self._ctx)
        out = self.modmul_records(modsel, to_records([cts[0].C % mod], width), inv, width)
        return Ciphertext(from_records(out, width)[0], level, MIXED)

    def NestedRandomizeWithBatch(self, cts: Sequence[Ciphertext], As: Sequence[int], Bs: Sequence[int]) -> List[Ciphertext]:
        """N x PublicKey.NestedRandomize (operations.go:96-118) with the randomness (a, b) supplied"""
        if any(c.Level != ENC_LEVEL_TWO for c in cts):
            raise ValueError("can only homomorphically randomize doubly encrypted values")          # :97-99
        cr, ar, br = to_records([c.C for c in cts], self.w_n3), to_records(As, self.w_n), to_records(Bs, self.w_n)
        out = np.empty(len(cts) * self.w_n3, dtype=np.uint8)
        check(lib.pgpu_nested_randomize_with(self._ctx, len(cts), _ptr(cr), _ptr(ar), _ptr(br), _ptr(out)), self._ctx)
        return [Ciphertext(c, ENC_LEVEL_TWO, REGULAR) for c in from_records(out, self.w_n3)]

    def _nested(self, fn, ct1s, ct2s):
        if any(c.Level != ENC_LEVEL_TWO for c in ct1s) or any(c.Level != ENC_LEVEL_ONE for c in ct2s):
            raise ValueError("can only homomorphically add an encrypted value to a doubly encrypted value")   # :122-124
        r1, r2 = to_records([c.C for c in ct1s], self.w_n3), to_records([c.C for c in ct2s], self.w_n2)
        out = np.empty(len(ct1s) * self.w_n3, dtype=np.uint8)
        check(fn(self._ctx, len(ct1s), _ptr(r1), _ptr(r2), _ptr(out)), self._ctx)
        return [Ciphertext(c, a.Level, a.EncMethod) for c, a in zip(from_records(out, self.w_n3), ct1s)]

    def NestedAddBatch(self, ct1s: Sequence[Ciphertext], ct2s: Sequence[Ciphertext]) -> List[Ciphertext]:
        """N x PublicKey.NestedAdd (operations.go:121-127)"""
        return self._nested(lib.pgpu_nested_add, ct1s, ct2s)

    def NestedSubBatch(self, ct1s: Sequence[Ciphertext], ct2s: Sequence[Ciphertext]) -> List[Ciphertext]:
        """N x PublicKey.NestedSub (operations.go:130-140)"""
        return self._nested(lib.pgpu_nested_sub, ct1s, ct2s)

    def VerifyDDLEQProofBatch(self, ct1s: Sequence[Ciphertext], ct2s: Sequence[Ciphertext], proofs: Sequence[DDLEQProof]) -> List[bool]:
        """N x PublicKey.VerifyDDLEQProof (ddleq.go:44-53); all proofs of a batch have the same number of instances"""
        if not proofs:
            return []
        secpar = len(proofs[0].Instances)
        if secpar == 0:
            return [True] * len(proofs)                                                               # empty loop, :46-51
        if any(len(p.Instances) != secpar for p in proofs):
            raise ValueError("VerifyDDLEQProofBatch: one secpar per batch")
        inst = [i for p in proofs for i in p.Instances]
        c1, c2 = to_records([c.C for c in ct1s], self.w_n3), to_records([c.C for c in ct2s], self.w_n3)
        x, y = to_records([i.X for i in inst], self.w_n), to_records([i.Y for i in inst], self.w_n)
        al, e, f = (to_records([i.Alpha for i in inst], self.w_n3), to_records([i.E for i in inst], self.w_n2),
                    to_records([i.F for i in inst], self.w_n3))
        ok = self.verify_ddleq_records(len(proofs), secpar, c1, c2, x, y, al, e, f)
        return [bool(ok[i * secpar:(i + 1) * secpar].all()) for i in range(len(proofs))]

    def verify_ddleq_records(self, count: int, secpar: int, c1, c2, x, y, al, e, f) -> np.ndarray:
        """record-level VerifyDDLEQProofBatch: count statements (n3-width ct1, ct2) x secpar instances (n-width x, y,
        n3-width alpha, n2-width e, n3-width f) -> one verdict byte per instance"""
        ok = np.zeros(count * secpar, dtype=np.uint8)
        check(lib.pgpu_ddleq_verify(self._ctx, count, secpar, _ptr(c1), _ptr(c2), _ptr(x), _ptr(y), _ptr(al), _ptr(e), _ptr(f),
                                    _ptr(ok)), self._ctx)
        return ok

    # -- device-resident batches: enqueued on the context's stream, NOT synchronised; Sync() waits for them ---------------
    def NewDeviceBuffer(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self._ctx, nbytes)

    def Sync(self) -> None:
        check(lib.pgpu_ctx_sync(self._ctx), self._ctx)

    def EncryptWithRDev(self, count: int, m: DeviceBuffer, r: DeviceBuffer, c: DeviceBuffer) -> None:
        """c[i] = EncryptWithR(m[i], r[i]) (paillier.go:185-187): n-width m and r, n2-width c"""
        check(lib.pgpu_encrypt_with_r_dev(self._ctx, count, _need(m, count * self.w_n), _need(r, count * self.w_n), _need(c, count * self.w_n2)),
              self._ctx)

    def ConstMultDev(self, count: int, c: DeviceBuffer, k: DeviceBuffer, k_bytes: int, out: DeviceBuffer) -> None:
        """out[i] = ConstMult(c[i], k[i]) (operations.go:58-64): k = k_bytes-wide little-endian unsigned scalars (multiple of 4)"""
        check(lib.pgpu_const_mult_dev(self._ctx, count, _need(c, count * self.w_n2), _need(k, count * k_bytes), k_bytes,
                                      _need(out, count * self.w_n2)), self._ctx)

    def AddPairsDev(self, count: int, a: DeviceBuffer, b: DeviceBuffer, out: DeviceBuffer) -> None:
        """out[i] = Add(a[i], b[i]) (operations.go:11-29); out may be a or b"""
        check(lib.pgpu_add_pairs_dev(self._ctx, count, _need(a, count * self.w_n2), _need(b, count * self.w_n2), _need(out, count * self.w_n2)),
              self._ctx)

    def AddReduceDev(self, count: int, c: DeviceBuffer, out: DeviceBuffer) -> None:
        """out = Add(c[0], ..., c[count-1]) as one tree reduction"""
        check(lib.pgpu_add_reduce_dev(self._ctx, count, _need(c, count * self.w_n2), _need(out, self.w_n2)), self._ctx)

    # -- introspection -------------------------------------------------------
    def launch_count(self) -> int:
        v = C.c_uint64()
        check(lib.pgpu_ctx_launch_count(self._ctx, C.byref(v)), self._ctx)
        return v.value

    def kernel_shape(self, modsel: int) -> dict:
        """exponentiation kernel serving a modulus (0 n, 1 n^2, 2 n^3, 3 p^2/q^2, 4 p^3/q^3)"""
        t, l, f, g = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib.pgpu_ctx_kernel_shape(self._ctx, modsel, C.byref(t), C.byref(l), C.byref(f), C.byref(g)), self._ctx)
        return {"tpi": t.value, "limbs_per_lane": l.value, "fp64": bool(f.value), "resident_groups": g.value}

    def program_cost(self, what: int):
        s, q, m = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib.pgpu_ctx_program_cost(self._ctx, what, C.byref(s), C.byref(q), C.byref(m)), self._ctx)
        return s.value, q.value, m.value


class SecretKey(PublicKey):
    """paillier.go:59-62: the reference keeps Lambda = (p-1)(q-1) only; either form is accepted."""

    def __init__(self, N: int, Lambda: Optional[int] = None, p: Optional[int] = None, q: Optional[int] = None, device: int = 0,
                 H: Optional[int] = None, K: Optional[int] = None):
        super().__init__(N, device, H, K)
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

    def encrypt_with_r_records(self, m, r) -> np.ndarray:
        """EncryptWithR reached through the embedded PublicKey (paillier.go:29-34): same ciphertexts as
        PublicKey.encrypt_with_r_records, computed over p^2 and q^2 (pgpu_encrypt_with_r_sk)."""
        m = np.ascontiguousarray(m).view(np.uint8).reshape(-1)
        count = m.size // self.w_n
        m = _as_u8(m, count * self.w_n, "m")
        r = _as_u8(r, count * self.w_n, "r")
        out = np.empty(count * self.w_n2, dtype=np.uint8)
        check(lib.pgpu_encrypt_with_r_sk(self._ctx, count, _ptr(m), _ptr(r), _ptr(out)), self._ctx)
        return out

    def EncryptWithRAtLevelBatch(self, ms: Sequence[int], rs: Sequence[int], level: int) -> List[Ciphertext]:
        """EncryptWithRAtLevel reached through the embedded PublicKey: same ciphertexts, r^(n^s) over the prime powers"""
        wm, wc = self._level_widths(level)
        mr, rr = to_records(ms, wm), to_records(rs, self.w_n)
        out = np.empty(len(ms) * wc, dtype=np.uint8)
        check(lib.pgpu_encrypt_with_r_at_level_sk(self._ctx, level + 1, len(ms), _ptr(mr), _ptr(rr), _ptr(out)), self._ctx)
        return [Ciphertext(c, level, REGULAR) for c in from_records(out, wc)]

    def decrypt_records(self, c) -> np.ndarray:
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // self.w_n2
        c = _as_u8(c, count * self.w_n2, "c")
        out = np.empty(count * self.w_n, dtype=np.uint8)
        check(lib.pgpu_decrypt(self._ctx, count, _ptr(c), _ptr(out)), self._ctx)
        return out

    def DecryptDev(self, count: int, c: DeviceBuffer, m: DeviceBuffer) -> None:
        """m[i] = Decrypt(c[i]) (paillier.go:292-303, CRT over p^2, q^2) on device buffers; enqueued, see Sync()"""
        check(lib.pgpu_decrypt_dev(self._ctx, count, _need(c, count * self.w_n2), _need(m, count * self.w_n)), self._ctx)

    def DecryptBatch(self, cts: Sequence[Ciphertext]) -> List[int]:
        """N x SecretKey.Decrypt (paillier.go:292-303); one level per batch"""
        if not cts:
            return []
        level = cts[0].Level
        if any(c.Level != level for c in cts):
            raise ValueError("DecryptBatch: one encryption level per batch")
        if level == ENC_LEVEL_ONE:
            return from_records(self.decrypt_records(to_records([c.C for c in cts], self.w_n2)), self.w_n)
        wm, wc = self._level_widths(level)
        cr = to_records([c.C for c in cts], wc)
        out = np.empty(len(cts) * wm, dtype=np.uint8)
        check(lib.pgpu_decrypt_at_level(self._ctx, level + 1, len(cts), _ptr(cr), _ptr(out)), self._ctx)
        return from_records(out, wm)

    def DecryptNestedCiphertextLayerBatch(self, cts: Sequence[Ciphertext]) -> List[Ciphertext]:
        """N x SecretKey.DecryptNestedCiphertextLayer (paillier.go:360-372)"""
        if any(c.Level == ENC_LEVEL_ONE for c in cts):
            raise ValueError("no nested ciphertexts to recover")                                      # :362-364
        return [Ciphertext(v, ENC_LEVEL_ONE, MIXED) for v in self.DecryptBatch(cts)]

    def NestedDecryptBatch(self, cts: Sequence[Ciphertext]) -> List[int]:
        """N x SecretKey.NestedDecrypt (paillier.go:344-356): the inner value 0 decrypts to 0 (:350-354)"""
        inner = self.DecryptNestedCiphertextLayerBatch(cts)
        nz = [i for i, c in enumerate(inner) if c.C != 0]
        vals = self.DecryptBatch([inner[i] for i in nz])
        out = [0] * len(cts)
        for i, v in zip(nz, vals):
            out[i] = v
        return out

    def ExtractRandonnessBatch(self, cts: Sequence[Ciphertext]) -> List[int]:
        """N x SecretKey.ExtractRandonness (operations.go:75-91); one level per batch"""
        if not cts:
            return []
        level = cts[0].Level
        _, wc = self._level_widths(level)
        cr = to_records([c.C for c in cts], wc)
        out = np.empty(len(cts) * self.w_n, dtype=np.uint8)
        check(lib.pgpu_extract_randomness(self._ctx, level + 1, len(cts), _ptr(cr), _ptr(out)), self._ctx)
        return from_records(out, self.w_n)

    def ProveDDLEQBatch(self, secpar: int, ct1s: Sequence[Ciphertext], ct2s: Sequence[Ciphertext], As: Sequence[int], Bs: Sequence[int],
                        xs: Sequence[Sequence[int]], ys: Sequence[Sequence[int]]) -> List[DDLEQProof]:
        """N x SecretKey.ProveDDLEQ (ddleq.go:27-40); xs[i][j], ys[i][j] in Z*_n are the randomness of instance j of
        statement i (ddleq.go:71-79).  Raises PgpuError where the reference panics on wrong inputs (:67-69)."""
        count = len(ct1s)
        if secpar == 0 or count == 0:
            return [DDLEQProof([]) for _ in range(count)]
        fx = [x for row in xs for x in row]
        fy = [y for row in ys for y in row]
        if len(fx) != count * secpar or len(fy) != count * secpar:
            raise ValueError("ProveDDLEQBatch: secpar values of x and y per statement")
        c1, c2 = to_records([c.C for c in ct1s], self.w_n3), to_records([c.C for c in ct2s], self.w_n3)
        a, b = to_records(As, self.w_n), to_records(Bs, self.w_n)
        x, y = to_records(fx, self.w_n), to_records(fy, self.w_n)
        al, e, f = self.prove_ddleq_records(count, secpar, c1, c2, a, b, x, y)
        A, E, F = from_records(al, self.w_n3), from_records(e, self.w_n2), from_records(f, self.w_n3)
        return [DDLEQProof([DDLEQProofInstance(fx[k], fy[k], A[k], E[k], F[k]) for k in range(i * secpar, (i + 1) * secpar)])
                for i in range(count)]


    def prove_ddleq_records(self, count: int, secpar: int, c1, c2, a, b, x, y):
        """record-level ProveDDLEQBatch -> (alpha, e, f) records of count * secpar instances"""
        total = count * secpar
        al, e, f = (np.empty(total * self.w_n3, dtype=np.uint8), np.empty(total * self.w_n2, dtype=np.uint8),
                    np.empty(total * self.w_n3, dtype=np.uint8))
        check(lib.pgpu_ddleq_prove(self._ctx, count, secpar, _ptr(c1), _ptr(c2), _ptr(a), _ptr(b), _ptr(x), _ptr(y),
                                   _ptr(al), _ptr(e), _ptr(f)), self._ctx)
        return al, e, f


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
        # A value that does not fit its record cannot come from an honest prover (E is a SHA-256 digest, Z < 2^(8 w_z) for
        # every r < n^2; c and c_i are residues mod n^2 in the reference's exponentiations): the reference's VerifyProof
        # just returns false for such a proof, so it is answered without the GPU instead of raising.
        n2 = self.N ** 2
        fits = [0 <= p.E < (1 << 256) and 0 <= p.Z < (1 << (8 * self.w_z)) and p.C >= 0 and p.Decryption >= 0 for p in proofs]
        val = lambda p, f, v: v if f else 0
        ok = self.verify_proof_records(proofs[0].ID, to_records([val(p, f, p.C % n2) for p, f in zip(proofs, fits)], self.w_n2),
                                       to_records([val(p, f, p.Decryption % n2) for p, f in zip(proofs, fits)], self.w_n2),
                                       to_records([val(p, f, p.E) for p, f in zip(proofs, fits)], 32),
                                       to_records([val(p, f, p.Z) for p, f in zip(proofs, fits)], self.w_z))
        return [bool(x) and f for x, f in zip(ok, fits)]

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

    def combine_verified_records(self, ids: Sequence[int], decs, ok):
        """pgpu_combine_verified: decs = len(ids) batches of n2-width records, ok = len(ids) * count verdict bytes (server
        major) -> (plaintext records, per-ciphertext flags: 0 where fewer than Threshold shares verified)"""
        decs = np.ascontiguousarray(decs).view(np.uint8).reshape(-1)
        ok = np.ascontiguousarray(ok, dtype=np.uint8).reshape(-1)
        k = len(ids)
        count = decs.size // (self.w_n2 * k) if k else 0
        out = np.zeros(count * self.w_n, dtype=np.uint8)
        item_ok = np.zeros(count, dtype=np.uint8)
        idarr = (C.c_int * max(k, 1))(*ids)
        rc = lib.pgpu_combine_verified(self._ctx, count, k, idarr, _ptr(decs) if decs.size else None, _ptr(ok) if ok.size else None,
                                       _ptr(out) if count else None, _ptr(item_ok) if count else None)
        if rc not in (PGPU_OK, PGPU_ERR_THRESHOLD):
            check(rc, self._ctx)
        return out, item_ok

    def CombinePartialDecryptionsZKPBatch(self, shares: Sequence[Sequence[PartialDecryptionZKP]], strict: bool = True):
        """N x CombinePartialDecryptionsZKP (thresholdkey.go:164-172): shares[j] is server j's batch, all batches in the same
        ciphertext order.  As in the reference the proofs filter PER CIPHERTEXT: ciphertext i is combined from the servers
        whose proof for i verifies.  Where fewer than Threshold remain the reference returns "Threshold not meet" for that
        ciphertext: with strict (default) the call raises PgpuError(PGPU_ERR_THRESHOLD) if that happens to any of them, with
        strict=False those positions hold None and the others their plaintext."""
        if not shares:
            raise PgpuError(PGPU_ERR_THRESHOLD, "Threshold not meet")
        count = len(shares[0])
        if any(len(s) != count for s in shares):
            raise ValueError("CombinePartialDecryptionsZKPBatch: one proof per ciphertext and server")
        ids = [s[0].ID if count else 0 for s in shares]
        if len(set(ids)) != len(ids) and count:
            raise PgpuError(PGPU_ERR_THRESHOLD, "two shares has been created by the same server")
        ok = np.array([self.VerifyProofBatch(s) for s in shares], dtype=np.uint8).reshape(-1)
        flat = [p.Decryption for s in shares for p in s]
        out, item_ok = self.combine_verified_records(ids, to_records(flat, self.w_n2), ok)
        vals = from_records(out, self.w_n)
        if strict and not item_ok.all():
            raise PgpuError(PGPU_ERR_THRESHOLD, f"Threshold not meet for {int(count - item_ok.sum())} of {count} ciphertexts")
        return [v if f else None for v, f in zip(vals, item_ok)]


    def VerifyDecryptionBatch(self, encryptedMessages: Sequence[int], decryptedMessages: Sequence[int],
                              shares: Sequence[Sequence[PartialDecryptionZKP]]) -> None:
        """N x ThresholdPublicKey.VerifyDecryption (thresholdkey.go:175-189); raises ValueError with the reference's
        error strings, or the combine error (PgpuError) when too few valid shares remain."""
        for s in shares:
            if [p.C for p in s] != list(encryptedMessages):
                raise ValueError("The encrypted message is not the same than the one in the shares")
        if self.CombinePartialDecryptionsZKPBatch(shares) != list(decryptedMessages):
            raise ValueError("The decrypted message is not the same than the one in the shares")


class ThresholdSecretKey(ThresholdPublicKey):
    """thresholdkey.go:38-42"""

    def __init__(self, N: int, TotalNumberOfDecryptionServers: int, Threshold: int, VerificationKey: int,
                 VerificationKeys: Sequence[int], ID: int, Share: int, device: int = 0):
        super().__init__(N, TotalNumberOfDecryptionServers, Threshold, VerificationKey, VerificationKeys,
                         device, _id=ID, _share=Share)
        self.ID = ID
        self.Share = Share

    def PublicKey(self) -> ThresholdPublicKey:
        """ThresholdSecretKey.PublicKey (thresholdkey.go:213-222): the key without ID and Share, on a context of its own"""
        return ThresholdPublicKey(self.N, self.TotalNumberOfDecryptionServers, self.Threshold, self.VerificationKey,
                                  self.VerificationKeys, device=self.device)

    def VerifyPartialDecryption(self, count: int = 1, rand=None) -> None:
        """ThresholdSecretKey.VerifyPartialDecryption (thresholdkey.go:258-275): encrypt a random m < n, prove its partial
        decryption, verify the proof; `count` such checks run as one batch.  Raises ValueError("Invalid share")."""
        import secrets
        below = (lambda b: rand.randrange(b)) if rand is not None else secrets.randbelow
        cts = self.EncryptBatch([below(self.N) for _ in range(count)], rand)
        proofs = self.PartialDecryptionWithZKPBatch([c.C for c in cts], [below(self.N ** 2) for _ in range(count)])
        if not all(self.VerifyProofBatch(proofs)):
            raise ValueError("Invalid share")

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

    def PartialDecryptDev(self, count: int, c: DeviceBuffer, out: DeviceBuffer) -> None:
        """out[i] = PartialDecrypt(c[i]) (thresholdkey.go:192-201) on device buffers of n2-width records; enqueued, see Sync()"""
        check(lib.pgpu_partial_decrypt_dev(self._ctx, count, _need(c, count * self.w_n2), _need(out, count * self.w_n2)), self._ctx)

    def PartialDecryptBatch(self, cs: Sequence[int]) -> List[PartialDecryption]:
        """N x ThresholdSecretKey.PartialDecrypt (thresholdkey.go:192-201)"""
        out = self.partial_decrypt_records(to_records(cs, self.w_n2))
        return [PartialDecryption(self.ID, v) for v in from_records(out, self.w_n2)]
