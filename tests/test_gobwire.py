"""Ciphertext wire format (paillier.go:374-401, paillier_test.go:140-156 TestToFromBytes).

PARITY UNPINNED: encoding/gob and github.com/ncw/gmp are not in /root/reference and no Go toolchain is available, so
the expected bytes here are assembled by hand from the encoding/gob specification.  The specification's own worked
example (type Point struct{X, Y int}; Point{22, 33}) pins the primitives."""
import random

import pytest

from paillier_b200 import gobwire as W


def test_primitives_match_the_gob_specification_example():
    # encoding/gob package documentation, "Encoding Details": the 31-byte type descriptor and 7-byte value of Point{22, 33}
    struct_t = b"\x01" + W._common("Point", 65) + b"\x01" + W._uint(2) + W._field("X", 2) + W._field("Y", 2) + b"\x00"
    typedef = W._message(W._int(-65) + b"\x03" + struct_t + b"\x00")
    assert typedef.hex() == "1fff81030101" + "05506f696e74" + "01ff8200" + "0102" + "010158010400" + "010159010400" + "0000"
    value = W._message(W._int(65) + b"\x01" + W._int(22) + b"\x01" + W._int(33) + b"\x00")
    assert value.hex() == "07ff82012c014200"
    # unsigned: < 128 one byte, else negated byte count + big-endian bytes; signed: sign in bit 0
    assert W._uint(7) == b"\x07" and W._uint(256) == b"\xfe\x01\x00" and W._uint(255) == b"\xff\xff"
    assert W._int(-129) == b"\xfe\x01\x01" and W._int(-1) == b"\x01" and W._int(0) == b"\x00"


def test_gmp_int_gob_layout():
    # ncw/gmp Int.GobEncode = math/big's: version 1 << 1 | sign, big-endian magnitude
    assert W.gmp_gob_encode(0) == b"\x02"
    assert W.gmp_gob_encode(1) == b"\x02\x01"
    assert W.gmp_gob_encode(-255) == b"\x03\xff"
    assert W.gmp_gob_encode(0x0102030405) == b"\x02\x01\x02\x03\x04\x05"
    for v in (0, 1, -1, 2 ** 64, -(2 ** 4096) + 12345):
        assert W.gmp_gob_decode(W.gmp_gob_encode(v)) == v
    with pytest.raises(W.GobError):
        W.gmp_gob_decode(b"\x04\x01")           # version 2


def test_ciphertext_stream_layout():
    # Ciphertext{C: 0x1234, Level: EncLevelOne (0, not sent), EncMethod: RegularEncryption (0, not sent)}
    b = W.encode_ciphertext(0x1234, 0, 0)
    t1 = ("ff81" "03" "01" "01" "0a" + b"Ciphertext".hex() + "01" "ff82" "00" "01" "03"
          "01" "01" "43" "01" "ff84" "00"
          "01" "05" + b"Level".hex() + "01" "04" "00"
          "01" "09" + b"EncMethod".hex() + "01" "04" "00"
          "00" "00")
    t2 = "ff83" "05" "01" "01" "03" + b"Int".hex() + "01" "ff84" "00" "00" "00"
    v = "ff82" "01" "03" "021234" "00"
    expect = bytes([len(t1) // 2]) + bytes.fromhex(t1) + bytes([len(t2) // 2]) + bytes.fromhex(t2) + bytes([len(v) // 2]) + bytes.fromhex(v)
    assert b == expect
    # Level two, alternative encryption: both fields sent with delta 1
    b = W.encode_ciphertext(5, 1, 1)
    assert b.endswith(bytes.fromhex("0b" "ff82" "01" "02" "0205" "01" "02" "01" "02" "00"))
    # only EncMethod set: delta 2 from C
    b = W.encode_ciphertext(5, 0, 2)
    assert b.endswith(bytes.fromhex("09" "ff82" "01" "02" "0205" "02" "04" "00"))


def test_round_trip_like_TestToFromBytes():
    rnd = random.Random(20)
    for bits in (1, 7, 8, 64, 127, 128, 1016, 1024, 4096, 6144):
        for level, method in ((0, 0), (1, 0), (0, 1), (1, 2)):
            c = rnd.getrandbits(bits) | (1 << (bits - 1))
            assert W.decode_ciphertext(W.encode_ciphertext(c, level, method)) == (c, level, method)
    assert W.decode_ciphertext(W.encode_ciphertext(0, 0, 0)) == (0, 0, 0)


def test_decoder_accepts_other_type_ids_and_rejects_garbage():
    # a sender whose process had registered other gob types first
    b = W.encode_ciphertext(123456789, 1, 1, struct_id=71)
    assert W.decode_ciphertext(b) == (123456789, 1, 1)
    with pytest.raises(W.GobError, match="no data provided"):
        W.decode_ciphertext(b"")
    good = W.encode_ciphertext(2 ** 300 + 1, 0, 0)
    for cut in (1, 10, len(good) // 2, len(good) - 1):
        with pytest.raises(W.GobError):
            W.decode_ciphertext(good[:cut])
    with pytest.raises(W.GobError):
        W.decode_ciphertext(good[good.index(b"\xff\x83") - 1:])     # value without its struct definition


def test_api_methods():
    from paillier_b200.api import Ciphertext, PublicKey
    ct = Ciphertext(0xDEADBEEF, 1, 2)
    back = PublicKey.NewCiphertextFromBytes(None, ct.Bytes())
    assert back == ct
