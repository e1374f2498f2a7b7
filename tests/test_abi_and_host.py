"""CPU-only checks: the C-ABI library loads, exports every symbol include/pgpu.h declares,
fails loudly without a GPU, and its host-side big-integer code (key setup) matches Python."""
import ctypes as C
import os
import random
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from paillier_b200 import _lib
    return _lib


def test_header_symbols_are_exported_and_bound():
    L = _lib()
    header = open(os.path.join(ROOT, "include", "pgpu.h")).read()
    declared = set(re.findall(r"\b(pgpu_[a-z0-9_]+)\s*\(", header))
    declared.discard("pgpu_ctx")
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L.lib, name), f"{name} declared in pgpu.h but not exported"
    assert declared == set(L.SIGNATURES), (declared ^ set(L.SIGNATURES))
    assert L.lib.pgpu_version() == 1


def test_go_binding_and_cpp_mirror_use_declared_symbols_only():
    """go/*.go (uncompiled here) and include/paillier_b200.hpp may only call what pgpu.h declares, with the right arity."""
    header = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "pgpu.h")).read(), flags=re.S)
    arity = {}
    for name, args in re.findall(r"\b(pgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        arity[name] = 0 if args.strip() in ("", "void") else args.count(",") + 1

    def calls(text, prefix):
        out = []
        for m in re.finditer(prefix + r"(pgpu_[a-z0-9_]+)\s*\(", text):
            depth, i, n_args, any_arg = 1, m.end(), 0, False
            while depth:
                ch = text[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    n_args += 1
                if depth and not ch.isspace():
                    any_arg = True
                i += 1
            out.append((m.group(1), n_args + 1 if any_arg else 0))
        return out

    go_dir = os.path.join(ROOT, "go")
    used = []
    for f in sorted(os.listdir(go_dir)):
        if f.endswith(".go"):
            used += calls(open(os.path.join(go_dir, f)).read(), r"C\.")
    assert len({n for n, _ in used}) >= 25
    hpp = open(os.path.join(ROOT, "include", "paillier_b200.hpp")).read()
    used += calls(hpp, r"(?<![A-Za-z_])")
    for name, n in used:
        assert name in arity, f"{name} is not declared in pgpu.h"
        assert n == arity[name], f"{name}: called with {n} arguments, declared with {arity[name]}"


def test_no_torch_types_in_abi():
    header = open(os.path.join(ROOT, "include", "pgpu.h")).read()
    assert "torch" not in header.lower() and "at::" not in header


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib()
    ctx = C.c_void_p()
    n = (101 * 103).to_bytes(2, "big")
    rc = L.lib.pgpu_ctx_create(C.byref(ctx), 0, n, len(n))
    assert rc == L.PGPU_ERR_CUDA
    assert b"CUDA" in L.lib.pgpu_last_error(None) or b"device" in L.lib.pgpu_last_error(None)
    from paillier_b200.api import PublicKey
    with pytest.raises(L.PgpuError):
        PublicKey(101 * 103)


def _bn(op, a, b=0, m=0):
    L = _lib()
    be = lambda x: x.to_bytes((x.bit_length() + 7) // 8, "big")
    out = C.create_string_buffer(4096)
    n = C.c_size_t(4096)
    ab, bb, mb = be(a), be(b), be(m)
    rc = L.lib.pgpu_selftest_bn(op, ab, len(ab), bb, len(bb), mb, len(mb), out, C.byref(n))
    if rc != 0:
        return None
    return int.from_bytes(out.raw[:n.value], "big")


def test_host_bignum_against_python():
    rnd = random.Random(20260101)
    for it in range(300):
        abits = rnd.choice([1, 31, 32, 33, 64, 100, 1024, 2048, 4096, 6144])
        bbits = rnd.choice([1, 17, 32, 33, 63, 64, 65, 512, 1024, 2048])
        a = rnd.getrandbits(abits) | (1 << (abits - 1))
        b = rnd.getrandbits(bbits) | (1 << (bbits - 1))
        assert _bn(0, a, b) == a * b
        assert _bn(1, a, b) == a // b
        assert _bn(2, a, b) == a % b
        m = rnd.getrandbits(bbits + 7) | 1 | (1 << (bbits + 6))
        inv = _bn(3, a, 0, m)
        import math
        if math.gcd(a, m) == 1:
            assert inv == pow(a, -1, m)
        else:
            assert inv is None
        if it % 10 == 0:
            e = rnd.getrandbits(70)
            assert _bn(4, a, e, m) == pow(a, e, m)
        assert _bn(5, a) == math.isqrt(a)
    # Knuth D corner: qhat overestimate / add-back paths
    for a, b in [((1 << 128) - 1, (1 << 64) + 1), ((1 << 96), (1 << 64) - 1), (0x7fffffff800000010000000000000000, 0x800000008000000200000005),
                 ((1 << 4096) - 1, (1 << 2048) - 1), (3 * ((1 << 64) - 1) ** 2, (1 << 64) - 1)]:
        assert _bn(1, a, b) == a // b and _bn(2, a, b) == a % b


def test_every_export_survives_null_and_zero_arguments():
    """No abort across the ABI (SURVEY.md 8b "Errors"): every exported function called with a null context, null pointers and
    zero counts returns (an error code where it has one) instead of crashing.  Runs in a child process so that a crash is a
    test failure, not the end of the test session."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from paillier_b200 import _lib as L
ints = (C.c_int, C.c_uint, C.c_size_t, C.c_uint32, C.c_uint64)
for name, (res, args) in L.SIGNATURES.items():
    rc = getattr(L.lib, name)(*[0 if a in ints else None for a in args])
    print(name, rc, flush=True)
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    lines = dict(l.split(" ", 1) for l in r.stdout.strip().split("\n"))
    assert r.returncode == 0, f"crashed after {list(lines)[-1] if lines else 'start'}: {r.stderr[-500:]}"
    L = _lib()
    assert set(lines) == set(L.SIGNATURES)
    ok_with_nulls = {"pgpu_version", "pgpu_ctx_destroy", "pgpu_buf_free", "pgpu_host_free", "pgpu_multi_destroy",     # destroy(NULL) is a no-op
                     "pgpu_buf_size", "pgpu_multi_size", "pgpu_buf_ptr", "pgpu_last_error", "pgpu_primes_last_error", "pgpu_multi_last_error"}
    for name, rc in lines.items():
        if name not in ok_with_nulls:
            assert rc not in ("0", "None"), f"{name} accepted null arguments"


def test_host_bignum_structured_operands():
    """bn_host.hpp on operands built to hit carries, borrows and normalisation: all-ones, single bits, 2^k - small, limbs drawn
    from {0, 0xffffffff, 0x80000000, 1, random}, widths around the 32-bit and 64-bit limb boundaries up to 8191 bits."""
    import math
    rnd = random.Random(4242)

    def rv():
        bits = rnd.choice([0, 1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 255, 256, 1000, 2048, 4096, 6144, 8191])
        if bits == 0:
            return 0
        k = rnd.randrange(8)
        if k == 0:
            return (1 << bits) - 1
        if k == 1:
            return 1 << (bits - 1)
        if k == 2:
            return max(0, (1 << bits) - rnd.randrange(1, 4))
        if k == 3:
            v = 0
            for i in range(0, bits, 32):
                v |= rnd.choice([0, 0xffffffff, 0x80000000, 1, rnd.getrandbits(32)]) << i
            return v & ((1 << bits) - 1)
        return rnd.getrandbits(bits)

    for it in range(400):
        a, b = rv(), rv()
        assert _bn(0, a, b) == a * b
        if b:
            assert _bn(1, a, b) == a // b and _bn(2, a, b) == a % b
        m = rv() | 1
        if m > 1:
            want = pow(a, -1, m) if math.gcd(a, m) == 1 else None
            assert _bn(3, a, 0, m) == want
            if it % 7 == 0:
                e = rnd.getrandbits(rnd.choice([1, 8, 64, 130]))
                assert _bn(4, a, e, m) == pow(a, e, m)
        assert _bn(5, a) == math.isqrt(a)
