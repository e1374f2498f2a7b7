"""The kernels' multiplier SOURCES executed on the CPU (no GPU needed).

tests/cpp/mont_host_test.cpp includes paillier_b200/csrc/mont.cuh and mont52.cuh unchanged -- the same headers powm.cu compiles for
sm_100a -- on top of tests/cpp/cuda_host_shim.h, which turns a warp into 32 real threads (one per lane; shuffles, ballots and
__syncwarp are barriers; the dedicated squaring's shared memory is an array the threads share; fma_rz is fma() under
FE_TOWARDZERO; the carry flag of the mad.cc chains is a thread-local).  Every shape powm.cu builds runs Montgomery mul, sqr
(the shared-memory squaring where the product uses it), add and sub for every lane group of the warp, against Python integers.
tools/host_sanitize.sh runs the same binary under ThreadSanitizer: with one lane per thread a missing __syncwarp() in the
squaring's exchange is a real data race (profiles/r02_host_sanitizers.txt: none; a negative control with one __syncwarp removed
is reported)."""
import os
import random
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "mont_host_test")


def _shapes():
    src = open(os.path.join(ROOT, "paillier_b200", "csrc", "powm.cu")).read()
    s32 = src[src.index("#define PGPU_FOR_EACH_SHAPE(X)"):src.index("// FP64-pipe shapes")]
    s52 = src[src.index("#define PGPU_FOR_EACH_SHAPE52(X)"):src.index("cudaError_t vm_launch")]
    i32 = [(int(a), int(b)) for a, b in re.findall(r"X\((\d+),\s*(\d+)\)", s32)]
    i52 = [(int(a), int(b), int(c)) for a, b, c in re.findall(r"X\((\d+),\s*(\d+),\s*(\d+)\)", s52)]
    return i32, i52


@pytest.fixture(scope="module")
def exe():
    # PGPU_EMU_CXXFLAGS="-fsanitize=thread -g" turns the whole file into the ThreadSanitizer run of tools/host_sanitize.sh: a report
    # makes the binary exit with 66, which every test below treats as a failure
    extra = os.environ.get("PGPU_EMU_CXXFLAGS", "").split()
    subprocess.run(["g++", "-std=c++20", "-O1", "-frounding-math", "-pthread", "-Wno-unknown-pragmas"] + extra +
                   [os.path.join(ROOT, "tests", "cpp", "mont_host_test.cpp"), "-o", EXE], check=True, capture_output=True, text=True)
    return EXE


def _moduli(rnd, bits):
    yield rnd.getrandbits(bits) | 1 | (1 << (bits - 1))              # full width
    yield (1 << bits) - 1 - 2 * rnd.getrandbits(20)                   # just below 2^bits
    yield (1 << (bits - 1)) + 1 + 2 * rnd.getrandbits(8)              # just above 2^(bits-1)
    yield rnd.getrandbits(rnd.randrange(40, bits)) | 3                # short modulus in a wide record


def test_the_harness_covers_every_built_shape():
    i32, i52 = _shapes()
    src = open(os.path.join(ROOT, "tests", "cpp", "mont_host_test.cpp")).read()
    lists = src[src.index("#define SHAPES32(X)"):src.index("// the interpreter is instantiated for a subset")]
    h32 = {(int(a), int(b)) for a, b in re.findall(r"X\((\d+), (\d+), (?:true|false)\)", lists)}
    h52 = {(int(a), int(b), int(c)) for a, b, c in re.findall(r"X\((\d+), (\d+), (\d+)\)", lists)}
    assert set(i32) == h32 and set(i52) == h52


def test_integer_pipe_multiplier_source_on_emulated_warps(exe):
    i32, _ = _shapes()
    rnd = random.Random(32)
    lines, want = [], []
    for tpi, l in i32:
        nsm = int(tpi == 4 and l % 8 == 0 and l <= 16)                # SqrShape: these shapes square through shared memory
        bits, groups = 32 * tpi * l, 32 // tpi
        R = 1 << bits
        for n in _moduli(rnd, bits):
            pairs = [(rnd.randrange(n), rnd.randrange(n)) for _ in range(groups)]
            pairs[0] = (n - 1, n - 1)
            if groups > 1:
                pairs[1] = (0, rnd.randrange(n))
            full = groups - 1                                          # last group: a = R - 1 (mul allows a < R), mul only
            pairs[full] = (R - 1, n - 1)
            lines.append(f"m32 {tpi} {l} {nsm} {n:x} {groups} " + " ".join(f"{a:x} {b:x}" for a, b in pairs))
            ri = pow(R, -1, n)
            for g, (a, b) in enumerate(pairs):
                want.append((a * b * ri % n, None if g == full else a * a * ri % n, None if g == full else (a + b) % n,
                             None if g == full else (a - b) % n))
    r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = [[int(x, 16) for x in ln.split()] for ln in r.stdout.strip().split("\n")]
    assert len(out) == len(want)
    for got, exp in zip(out, want):
        assert all(e is None or g == e for g, e in zip(got, exp)), (got, exp)


def test_fp64_pipe_multiplier_source_on_emulated_warps(exe):
    _, i52 = _shapes()
    rnd = random.Random(52)
    lines, want = [], []
    for tpi, l, s32 in i52:
        bits, groups = 32 * s32, 32 // tpi
        R = 1 << (52 * tpi * l)
        for n in _moduli(rnd, bits):
            pairs = [(rnd.randrange(n), rnd.randrange(n)) for _ in range(groups)]
            pairs[0] = (n - 1, n - 1)
            full = groups - 1                                          # a full-width record (>= n) times a residue: still < 2n, mul only
            pairs[full] = ((1 << bits) - 1, n - 1)
            lines.append(f"m52 {tpi} {l} {s32} {n:x} {groups} " + " ".join(f"{a:x} {b:x}" for a, b in pairs))
            ri = pow(R, -1, n)
            for g, (a, b) in enumerate(pairs):
                want.append((a * b * ri % n, None if g == full else (a + b) % n, None if g == full else (a - b) % n))
    r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = [[int(x, 16) for x in ln.split()] for ln in r.stdout.strip().split("\n")]
    assert len(out) == len(want)
    for got, exp in zip(out, want):
        assert all(e is None or g == e for g, e in zip(got, exp)), (got, exp)


def test_fp64_pipe_short_records(exe):
    """Mont52::load_rec / store_rec with records narrower than the modulus: EncryptWithR reads n-width m and r into the n^2 shape
    (OP_LDI with in_limbs < S), n-width results are stored from it (out_limbs < S).  b = R mod n makes the product equal a."""
    _, i52 = _shapes()
    rnd = random.Random(53)
    lines, want = [], []
    for tpi, l, s32 in i52:
        groups = 32 // tpi
        R = 1 << (52 * tpi * l)
        n = rnd.getrandbits(32 * s32) | 1 | (1 << (32 * s32 - 1))
        for lim in (s32 // 2, 1, s32 - 1, s32 // 2 + 3):
            pairs = [(rnd.getrandbits(32 * lim), R % n) for _ in range(groups)]
            pairs[0] = ((1 << (32 * lim)) - 1, R % n)
            lines.append(f"m52s {tpi} {l} {s32} {lim} {n:x} {groups} " + " ".join(f"{a:x} {b:x}" for a, b in pairs))
            want += [a for a, _ in pairs]
    r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert [int(x, 16) for x in r.stdout.split()] == want


# ---- the interpreter's source (csrc/vm_run.cuh, what powm_vm / powm_vm52 wrap) running programs the LIBRARY compiled -----------

def _compile(kind, mod, base, shared=0, exps=(), exp_limbs=0, pre=1):
    """pgpu_selftest_program compiles (and interprets on the host's big integers); pgpu_selftest_last_program hands out the ops"""
    import ctypes as C
    from paillier_b200._lib import check, lib
    be = lambda x: x.to_bytes(max(1, (x.bit_length() + 7) // 8), "big")
    mb, bb, sb = be(mod), be(base), be(shared)
    k = len(exps)
    arr = (C.c_uint32 * max(1, k * exp_limbs))()
    for s, e in enumerate(exps):
        for i in range(exp_limbs):
            arr[s * exp_limbs + i] = (e >> (32 * i)) & 0xFFFFFFFF
    out = C.create_string_buffer(16 * len(mb))
    n_out, n_sqr, n_mul = C.c_uint32(), C.c_uint32(), C.c_uint32()
    check(lib.pgpu_selftest_program(kind, mb, len(mb), bb, len(bb), sb, len(sb) if shared else 0, C.cast(arr, C.c_void_p), exp_limbs, k, pre,
                                    C.cast(out, C.c_void_p), len(out.raw), C.byref(n_out), C.byref(n_sqr), C.byref(n_mul)))
    n_ops, tbl = C.c_size_t(), C.c_uint32()
    ops = (C.c_uint32 * 65536)()
    check(lib.pgpu_selftest_last_program(C.cast(ops, C.c_void_p), 65536, C.byref(n_ops), C.byref(tbl)))
    return list(ops[:n_ops.value]), tbl.value


VM_SHAPES = [("vm32", 4, 8, 32, 0), ("vm32", 4, 8, 32, 1), ("vm32", 4, 16, 64, 0), ("vm32", 8, 8, 64, 0), ("vm32", 4, 32, 128, 0),
             ("vm52", 4, 5, 32, 0), ("vm52", 8, 10, 128, 0)]


@pytest.mark.parametrize("kind,tpi,l,s32,flags", VM_SHAPES)
def test_interpreter_source_runs_compiled_programs(exe, kind, tpi, l, s32, flags):
    """EncryptWithR's sliding window over a shared exponent, ConstMult's fixed windows over per-item exponents, the combiner's
    shared-base multi-exponentiation and PartialDecrypt fused with the proof's power: compiled by the library, executed by
    vm_run on an emulated block (32 / TPI resident groups, one item more than that: a second round with idle groups), against pow().
    flags = 1 routes the squarings of the shapes with the shared-memory squaring through the general multiplier (PGPU_NO_SQR)."""
    rnd = random.Random(hash((kind, tpi, l, flags)) & 0xffff)
    bits = 32 * s32
    R = 1 << (32 * s32 if kind == "vm32" else 52 * tpi * l)
    n = rnd.getrandbits(bits) | 1 | (1 << (bits - 1))
    items = 32 // tpi + 1
    shape = f"{kind} {tpi} {l}" + (f" {s32}" if kind == "vm52" else "")
    bases = [rnd.randrange(n) for _ in range(items)]
    bases[0] = n - 1
    jobs, want = [], []

    def job(ops, tbl, exp_stride, exp_bits, exp_sub, outs, exps):
        head = f"{shape} {flags} {items} {s32} {exp_stride} {exp_bits} {exp_sub} {outs} {tbl} {n:x} {R % n:x} {R * R % n:x} {len(ops)} "
        jobs.append(head + " ".join(f"{o:x}" for o in ops) + " " + " ".join(f"{b:x} {e:x}" for b, e in zip(bases, exps)))

    # kind 0: base^e, e shared (r^n of EncryptWithR, c^(p-1) of Decrypt)
    e = rnd.getrandbits(61) | 1 << 60
    ops, tbl = _compile(0, n, 3, shared=e)
    job(ops, tbl, 0, 0, 0, 1, [0] * items)
    want += [[pow(b, e, n), 0] for b in bases]
    # kind 1: base^e_i, per-item exponents (ConstMult)
    es = [rnd.getrandbits(64) for _ in range(items)]
    es[0], es[-1] = 0, (1 << 64) - 1
    ops, tbl = _compile(1, n, 3, exps=[5], exp_limbs=2)
    job(ops, tbl, 2, 64, 0, 1, es)
    want += [[pow(b, x, n), 0] for b, x in zip(bases, es)]
    # kind 2: (base^4)^e_s for k = 3 exponents per item sharing one squaring chain (the combiner's VerifyProof)
    k = 3
    ek = [[rnd.getrandbits(64) for _ in range(k)] for _ in range(items)]
    ek[0] = [0, 1, (1 << 64) - 1]
    ops, tbl = _compile(2, n, 3, exps=[1, 2, 3], exp_limbs=2, pre=4)
    job(ops, tbl, 2 * k, 64, 2, k, [sum(x << (64 * s) for s, x in enumerate(row)) for row in ek])
    want += [[pow(pow(b, 4, n), x, n) for x in row] + [0] for b, row in zip(bases, ek)]
    # kind 3: c^e1 and (c^4)^r from one chain (PartialDecryptionWithZKP)
    e1 = 40320 * 2 * rnd.getrandbits(40)
    rs = [rnd.getrandbits(64) for _ in range(items)]
    rs[0] = (1 << 64) - 1
    ops, tbl = _compile(3, n, 3, shared=e1, exps=[7], exp_limbs=2)
    job(ops, tbl, 2, 64, 0, 1, rs)
    want += [[pow(b, e1, n), pow(pow(b, 4, n), x, n)] for b, x in zip(bases, rs)]
    r = subprocess.run([exe], input="\n".join(jobs) + "\n", capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [[int(x, 16) for x in ln.split()] for ln in r.stdout.strip().split("\n")]
    assert got == want
