"""The exponentiation programs the library compiles on the host, executed WITHOUT a GPU.

pgpu_selftest_program builds the same micro-programs (csrc/vm.h) that powm_vm runs -- the sliding-window schedule over a shared
exponent (EncryptWithR's r^n, Decrypt's c^(p-1), PartialDecrypt's c^(2*delta*s)), the fixed-window program over per-item exponents
(ConstMult, the proofs' powers), the right-to-left bucket programs of r02 (the combiner's shared-base multi-exponentiation and
PartialDecrypt fused with the proof's (c^4)^r) -- and interprets them op by op with plain modular arithmetic on the host's big
integers.  The results must equal pow(); the program costs must be what the design claims."""
import ctypes as C
import random

import pytest

from paillier_b200._lib import check, lib


def _be(x: int) -> bytes:
    return x.to_bytes(max(1, (x.bit_length() + 7) // 8), "big")


def run_program(kind, mod, base, shared=0, exps=(), exp_limbs=0, pre=1):
    mb, bb, sb = _be(mod), _be(base), _be(shared)
    k = len(exps)
    arr = (C.c_uint32 * max(1, k * exp_limbs))()
    for s, e in enumerate(exps):
        assert e < 1 << (32 * exp_limbs)
        for i in range(exp_limbs):
            arr[s * exp_limbs + i] = (e >> (32 * i)) & 0xFFFFFFFF
    out = C.create_string_buffer(16 * len(mb))
    n_out, n_sqr, n_mul = C.c_uint32(), C.c_uint32(), C.c_uint32()
    check(lib.pgpu_selftest_program(kind, mb, len(mb), bb, len(bb), sb, len(sb) if shared else 0, C.cast(arr, C.c_void_p), exp_limbs, k, pre,
                                    C.cast(out, C.c_void_p), len(out.raw), C.byref(n_out), C.byref(n_sqr), C.byref(n_mul)))
    w = len(mb)
    vals = [int.from_bytes(out.raw[j * w:(j + 1) * w], "big") for j in range(n_out.value)]
    return vals, n_sqr.value, n_mul.value


def _odd_modulus(rnd, bits):
    return rnd.getrandbits(bits) | (1 << (bits - 1)) | 1


@pytest.mark.parametrize("bits", [64, 521, 1024, 2048])
def test_sliding_window_over_a_shared_exponent(bits):
    rnd = random.Random(bits)
    n = _odd_modulus(rnd, 2 * bits if bits < 1024 else bits)
    for e in [0, 1, 2, 3, 255, 256, (1 << bits) - 1, 1 << (bits - 1), rnd.getrandbits(bits), rnd.getrandbits(bits) | 1 << (bits - 1)]:
        base = rnd.randrange(n)
        (got,), sq, mu = run_program(0, n, base, shared=e)
        assert got == pow(base, e, n)
        if e.bit_length() >= 512:           # the window keeps the multiplications far below one per bit
            assert sq <= e.bit_length() + 1 and mu <= 2 + (1 << 6) + e.bit_length() // 5


@pytest.mark.parametrize("limbs", [1, 2, 8, 33])
def test_fixed_windows_over_a_per_item_exponent(limbs):
    rnd = random.Random(limbs)
    n = _odd_modulus(rnd, 512)
    for e in [0, 1, (1 << (32 * limbs)) - 1, rnd.getrandbits(32 * limbs), rnd.getrandbits(32 * limbs - 7)]:
        base = rnd.randrange(n)
        (got,), _, _ = run_program(1, n, base, exps=[e], exp_limbs=limbs)
        assert got == pow(base, e, n)


@pytest.mark.parametrize("k,limbs,pre", [(1, 4, 1), (3, 9, 4), (8, 9, 4), (8, 52, 4), (5, 2, 2)])
def test_shared_base_multi_exponentiation(k, limbs, pre):
    # the combiner's (c^4)^Z for the k share-holders of one ciphertext: one squaring chain, one bucket set per exponent
    rnd = random.Random(100 * k + limbs)
    n = _odd_modulus(rnd, 768)
    base = rnd.randrange(n)
    exps = [rnd.getrandbits(32 * limbs) for _ in range(k)]
    exps[0] = 0
    if k > 2:
        exps[1] = (1 << (32 * limbs)) - 1
        exps[2] = 1
    got, sq, mu = run_program(2, n, base, exps=exps, exp_limbs=limbs, pre=pre)
    assert got == [pow(pow(base, pre, n), e, n) for e in exps]
    bits = 32 * limbs
    assert sq <= bits + 8                                          # ONE chain of squarings for all k exponents
    if bits >= 1024:
        assert mu < k * (bits / 5 + 140)                           # about bits/w + 2 * 2^w multiplications per exponent


@pytest.mark.parametrize("bits", [96, 1024, 3100])
def test_partial_decrypt_fused_with_the_proofs_power(bits):
    # c_i = c^(2*delta*s) and a = (c^4)^r from one squaring chain (thresholdkey.go:199, :241-242)
    rnd = random.Random(bits)
    n = _odd_modulus(rnd, 1024)
    limbs = (bits + 31) // 32
    for e1, r in [(rnd.getrandbits(bits) | 1 << (bits - 1), rnd.getrandbits(32 * limbs)), (rnd.getrandbits(bits + 15) | 1 << (bits + 14), 0),
                  (40320 * 2 * rnd.getrandbits(bits - 20), (1 << (32 * limbs)) - 1), (1, 1), (rnd.getrandbits(bits // 2), rnd.getrandbits(32 * limbs))]:
        c = rnd.randrange(n)
        got, sq, mu = run_program(3, n, c, shared=e1, exps=[r], exp_limbs=limbs)
        assert got == [pow(c, e1, n), pow(pow(c, 4, n), r, n)]
        top = max(e1.bit_length(), 32 * limbs)
        assert sq <= top + 8                                       # the two exponentiations share their squarings


def test_zero_and_one_bases():
    n = _odd_modulus(random.Random(1), 256)
    for base in (0, 1, n - 1):
        got, _, _ = run_program(2, n, base, exps=[5, 0, 1 << 40], exp_limbs=2, pre=4)
        assert got == [pow(pow(base, 4, n), e, n) for e in (5, 0, 1 << 40)]
        got, _, _ = run_program(3, n, base, shared=12345, exps=[77], exp_limbs=1)
        assert got == [pow(base, 12345, n), pow(pow(base, 4, n), 77, n)]
