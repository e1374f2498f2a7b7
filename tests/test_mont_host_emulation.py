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
    subprocess.run(["g++", "-std=c++20", "-O1", "-frounding-math", "-pthread", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpp", "mont_host_test.cpp"), "-o", EXE], check=True, capture_output=True, text=True)
    return EXE


def _moduli(rnd, bits):
    yield rnd.getrandbits(bits) | 1 | (1 << (bits - 1))              # full width
    yield (1 << bits) - 1 - 2 * rnd.getrandbits(20)                   # just below 2^bits
    yield (1 << (bits - 1)) + 1 + 2 * rnd.getrandbits(8)              # just above 2^(bits-1)
    yield rnd.getrandbits(rnd.randrange(40, bits)) | 3                # short modulus in a wide record


def test_the_harness_covers_every_built_shape():
    i32, i52 = _shapes()
    src = open(os.path.join(ROOT, "tests", "cpp", "mont_host_test.cpp")).read()
    h32 = {(int(a), int(b)) for a, b in re.findall(r"X\((\d+), (\d+), (?:true|false)\)", src)}
    h52 = {(int(a), int(b), int(c)) for a, b, c in re.findall(r"X\((\d+), (\d+), (\d+)\)", src)}
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
