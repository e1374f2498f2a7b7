"""Wire format of a ciphertext: `Ciphertext.Bytes()` / `PublicKey.NewCiphertextFromBytes`
(/root/reference/paillier.go:374-401), i.e. Go's encoding/gob stream of

    type Ciphertext struct { C *gmp.Int; Level EncryptionLevel; EncMethod EncryptionMethod }   (paillier.go:64-69)

with a fresh `gob.NewEncoder` per ciphertext.  Host-side marshalling (the reference does it on the host too); no
arithmetic.  The format is restated from the encoding/gob specification (Go standard library, package documentation
"Encoding Details") and from github.com/ncw/gmp `Int.GobEncode` (the same layout as math/big: one byte
`version<<1 | sign` with version 1, then the big-endian magnitude) -- neither is in /root/reference and there is no Go
toolchain in this image, so PARITY IS UNPINNED for this module: the tests check the stream against the layout worked
out by hand from the specification's own example, round trips, and decoding of streams with other type ids.

Stream produced by Bytes() in a process where Ciphertext is the first type gob sees (type ids 65 and 66):
  message 1  type definition -65: wireType{StructT: {CommonType{"Ciphertext", 65}, Field: [{"C", 66}, {"Level", int}, {"EncMethod", int}]}}
  message 2  type definition -66: wireType{GobEncoderT: {CommonType{"Int", 66}}}
  message 3  value of type 65: field deltas; C as a byte slice (GobEncode output); Level / EncMethod omitted when zero
Each message is preceded by its byte count; unsigned integers are one byte below 128, otherwise the negated byte count
followed by the big-endian bytes; signed integers are (i << 1) or (^i << 1) | 1 in an unsigned integer.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

_T_BOOL, _T_INT, _T_UINT, _T_FLOAT, _T_BYTES, _T_STRING = 1, 2, 3, 4, 5, 6
_FIRST_USER_ID = 65
_GMP_GOB_VERSION = 1


class GobError(ValueError):
    pass


# ---- primitives ---------------------------------------------------------------------------------------------------

def _uint(u: int) -> bytes:
    if u < 0:
        raise GobError("negative value for an unsigned field")
    if u < 128:
        return bytes([u])
    b = u.to_bytes((u.bit_length() + 7) // 8, "big")
    return bytes([256 - len(b)]) + b


def _int(i: int) -> bytes:
    return _uint((~i << 1) | 1 if i < 0 else i << 1)


def _string(s: str) -> bytes:
    b = s.encode()
    return _uint(len(b)) + b


def _message(body: bytes) -> bytes:
    return _uint(len(body)) + body


class _Reader:
    def __init__(self, data: bytes, pos: int = 0, end: int = -1):
        self.d, self.p, self.end = data, pos, len(data) if end < 0 else end

    def left(self) -> int:
        return self.end - self.p

    def take(self, n: int) -> bytes:
        if n < 0 or n > self.left():
            raise GobError("unexpected EOF")
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def uint(self) -> int:
        b = self.take(1)[0]
        if b < 128:
            return b
        n = 256 - b
        if n > 8:
            raise GobError("encoded unsigned integer out of range")
        return int.from_bytes(self.take(n), "big")

    def int(self) -> int:
        u = self.uint()
        return ~(u >> 1) if u & 1 else u >> 1

    def string(self) -> str:
        # a Go string is arbitrary bytes (gob does not validate UTF-8): a name that is not UTF-8 is just a name that matches nothing
        return self.take(self.uint()).decode("utf-8", "surrogateescape")


# ---- gmp.Int (ncw/gmp Int.GobEncode / GobDecode) -------------------------------------------------------------------

def gmp_gob_encode(x: int) -> bytes:
    mag = abs(x)
    return bytes([_GMP_GOB_VERSION << 1 | (1 if x < 0 else 0)]) + mag.to_bytes((mag.bit_length() + 7) // 8, "big")


def gmp_gob_decode(b: bytes) -> int:
    if len(b) == 0:
        return 0                              # a nil *Int is sent as an empty slice by later versions of math/big
    if b[0] >> 1 != _GMP_GOB_VERSION:
        raise GobError(f"Int.GobDecode: encoding version {b[0] >> 1} not supported")
    v = int.from_bytes(b[1:], "big")
    return -v if b[0] & 1 else v


# ---- encoder -------------------------------------------------------------------------------------------------------

def _common(name: str, type_id: int) -> bytes:
    # CommonType{Name, Id}: field 0 then field 1, end of struct
    return b"\x01" + _string(name) + b"\x01" + _int(type_id) + b"\x00"


def _field(name: str, type_id: int) -> bytes:
    return b"\x01" + _string(name) + b"\x01" + _int(type_id) + b"\x00"


def _type_definitions(struct_id: int, int_id: int) -> bytes:
    fields = [("C", int_id), ("Level", _T_INT), ("EncMethod", _T_INT)]
    struct_t = (b"\x01" + _common("Ciphertext", struct_id) +                     # structType.CommonType
                b"\x01" + _uint(len(fields)) + b"".join(_field(n, t) for n, t in fields) +   # structType.Field
                b"\x00")
    msg1 = _message(_int(-struct_id) + b"\x03" + struct_t + b"\x00")            # wireType.StructT is field 2: delta 3
    enc_t = b"\x01" + _common("Int", int_id) + b"\x00"                            # gobEncoderType{CommonType}
    msg2 = _message(_int(-int_id) + b"\x05" + enc_t + b"\x00")                   # wireType.GobEncoderT is field 4: delta 5
    return msg1 + msg2


def encode_ciphertext(c: int, level: int, method: int, struct_id: int = _FIRST_USER_ID) -> bytes:
    """Ciphertext.Bytes() (paillier.go:392-401)."""
    int_id = struct_id + 1
    body = _int(struct_id)
    g = gmp_gob_encode(c)
    body += b"\x01" + _uint(len(g)) + g
    last = 0
    for idx, v in ((1, level), (2, method)):
        if v != 0:                                                                # zero-valued fields are not sent
            body += _uint(idx - last) + _int(v)
            last = idx
    body += b"\x00"
    return _type_definitions(struct_id, int_id) + _message(body)


# ---- decoder -------------------------------------------------------------------------------------------------------

def _read_common(r: _Reader) -> Tuple[str, int]:
    name, tid, f = "", 0, -1
    while True:
        d = r.uint()
        if d == 0:
            return name, tid
        f += d
        if f == 0:
            name = r.string()
        elif f == 1:
            tid = r.int()
        else:
            raise GobError("CommonType: unknown field")


def _read_struct_type(r: _Reader) -> List[Tuple[str, int]]:
    fields: List[Tuple[str, int]] = []
    f = -1
    while True:
        d = r.uint()
        if d == 0:
            return fields
        f += d
        if f == 0:
            _read_common(r)
        elif f == 1:
            for _ in range(r.uint()):
                name, tid, g = "", 0, -1
                while True:
                    dd = r.uint()
                    if dd == 0:
                        break
                    g += dd
                    if g == 0:
                        name = r.string()
                    elif g == 1:
                        tid = r.int()
                    else:
                        raise GobError("fieldType: unknown field")
                fields.append((name, tid))
        else:
            raise GobError("structType: unknown field")


def _read_wire_type(r: _Reader):
    """-> ("struct", [(name, id)...]) | ("gobenc", None)"""
    d = r.uint()
    kind = d - 1
    if kind == 2:
        out = ("struct", _read_struct_type(r))
    elif kind in (4, 5, 6):                      # GobEncoderT, BinaryMarshalerT, TextMarshalerT: {CommonType}
        f = r.uint()
        if f == 1:
            _read_common(r)
            if r.uint() != 0:
                raise GobError("gobEncoderType: unknown field")
        elif f != 0:
            raise GobError("gobEncoderType: unknown field")
        out = ("gobenc", None)
    else:
        raise GobError("type definition not used by a Ciphertext stream")
    if r.uint() != 0:
        raise GobError("wireType: more than one kind set")
    return out


def decode_ciphertext(data: bytes) -> Tuple[int, int, int]:
    """PublicKey.NewCiphertextFromBytes (paillier.go:374-390) -> (C, Level, EncMethod).
    Fields are matched by name as gob does; type ids are whatever the sender's process assigned."""
    if len(data) == 0:
        raise GobError("no data provided")                   # paillier.go:377-379
    types: Dict[int, tuple] = {}
    top = _Reader(bytes(data))
    while top.left() > 0:
        n = top.uint()
        r = _Reader(top.d, top.p, top.p + n)
        top.take(n)
        tid = r.int()
        if tid < 0:
            if -tid in types:
                raise GobError("duplicate type received")
            types[-tid] = _read_wire_type(r)
            continue
        t = types.get(tid)
        if t is None or t[0] != "struct":
            raise GobError("gob: type mismatch: no fields matched compiling decoder for Ciphertext")
        c, level, method, f = None, 0, 0, -1
        fields = t[1]
        while True:
            d = r.uint()
            if d == 0:
                break
            f += d
            if f >= len(fields):
                raise GobError("gob: field number out of range")
            name, ftid = fields[f]
            kind = types.get(ftid, (None,))[0] if ftid >= 64 else ftid
            if kind == "gobenc" or kind in (_T_BYTES, _T_STRING):
                raw = r.take(r.uint())
                if name == "C":
                    if kind != "gobenc":
                        raise GobError("gob: wrong type for field C")
                    c = gmp_gob_decode(raw)
            elif kind == _T_INT:
                v = r.int()
                if name == "Level":
                    level = v
                elif name == "EncMethod":
                    method = v
            elif kind in (_T_UINT, _T_BOOL, _T_FLOAT):
                r.uint()
                if name in ("Level", "EncMethod"):
                    raise GobError(f"gob: wrong type for field {name}")
            else:
                raise GobError("gob: cannot skip a field of this type")
        if r.left() != 0:
            raise GobError("gob: extra data in message")
        return (0 if c is None else c), level, method
    raise GobError("unexpected EOF")
