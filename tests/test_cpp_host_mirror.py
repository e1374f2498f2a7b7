"""The C++ host mirror (include/paillier_b200.hpp): it must compile and link against libpaillier_b200.so on any box
(CPU test) and reproduce the golden vectors through the GPU (gpu test)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")
V = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
I = lambda s: int(s, 16)


def _build():
    libdir = os.path.join(ROOT, "paillier_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
           "-o", EXE, "-L", libdir, "-lpaillier_b200", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_host_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


def _input():
    lines = []
    for name in ("paillier_64", "paillier_2048"):
        c = V["cases"][name]
        p, q = I(c["p"]), I(c["q"])
        e = c["encrypt"]
        lines += ["paillier", hex(p * q), hex((p - 1) * (q - 1)), "4"]
        for i in range(4):
            lines += [e["m"][i], e["r"][i], e["c"][i], c["const_mult"]["k"][i], c["const_mult"]["c"][i]]
        # add_all over the first 4 ciphertexts
        acc = 1
        for x in e["c"][:4]:
            acc = acc * I(x) % (p * q) ** 2
        lines.append(hex(acc))
    for name in ("threshold_512",):
        c = V["cases"][name]
        n = I(c["p"]) * I(c["q"])
        lines += ["threshold", hex(n), str(c["l"]), str(c["w"]), c["V"]] + c["vi"]
        for i in range(c["w"]):
            lines += [str(i + 1), c["shares"][i]]
        lines += [str(len(c["c"]))] + c["c"] + c["m"] + c["zkp_r"]
    for name in ("paillier_64", "paillier_2048"):
        c = V["cases"][name]
        p, q = I(c["p"]), I(c["q"])
        e1, e2, a, d = c["encrypt"], c["encrypt_level2"], c["alt_encrypt"], c["ddleq"]
        count = min(len(e2["m"]), len(e1["m"]))
        lines += ["level2", hex(p * q), hex((p - 1) * (q - 1)), c["H"], str(I(c["K"]).bit_length() - 1), str(count)]
        for i in range(count):
            lines += [e1["m"][i], e2["m"][i], e2["r"][i], e2["c"][i]]
        lines.append(str(len(a["r"])))
        for i in range(len(a["r"])):
            lines += [a["r"][i], a["c_level1"][i], a["c_level2"][i]]
        lines += [d["ct1"], d["ct2"], d["a"], d["b"], str(len(d["x"]))]
        for i in range(len(d["x"])):
            lines += [d["x"][i], d["y"][i], d["alpha"][i], d["e"][i], d["f"][i]]
    for bits, d in V["safe_prime"].items():
        lines += ["safeprime", bits, str(len(d["raw"]))]
        for raw, qv, ok in zip(d["raw"], d["q"], d["ok"]):
            lines += [raw, qv, str(int(ok))]
    return "\n".join(lines) + "\n"


@pytest.mark.gpu
def test_host_mirror_reproduces_golden_vectors():
    _build()
    r = subprocess.run([EXE], input=_input(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "host mirror ok" in r.stdout


def test_cpp_gob_wire_format_matches_python_mirror():
    """Ciphertext.Bytes / NewCiphertextFromBytes of the C++ mirror against paillier_b200/gobwire.py (no GPU needed)."""
    import random
    from paillier_b200 import gobwire as W
    libdir = os.path.join(ROOT, "paillier_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "gob_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "gob_test.cpp"),
                    "-o", exe, "-L", libdir, "-lpaillier_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    rnd = random.Random(8)
    cases = [(0, 0, 0), (1, 0, 0), (0x1234, 1, 1), (127, 0, 2), (128, 1, 0)]
    cases += [(rnd.getrandbits(b) | 1 << (b - 1), rnd.randrange(2), rnd.randrange(3)) for b in (64, 1016, 1024, 4096, 6144)]
    lines = [f"enc {c:x} {l} {m}" for c, l, m in cases]
    lines += ["dec " + W.encode_ciphertext(c, l, m, struct_id=65 + i).hex() for i, (c, l, m) in enumerate(cases)]
    good = W.encode_ciphertext(2 ** 200 + 5, 0, 0)
    lines += ["dec -", "dec " + good[:-3].hex(), "dec " + good[good.index(b"\xff\x83") - 1:].hex()]
    r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    out = r.stdout.split("\n")
    k = len(cases)
    assert out[:k] == [W.encode_ciphertext(c, l, m).hex() for c, l, m in cases]
    assert out[k:2 * k] == [f"{hex(c)} {l} {m}" for c, l, m in cases]
    assert out[2 * k] == "error no data provided"
    assert out[2 * k + 1].startswith("error ") and out[2 * k + 2].startswith("error ")


CALLERS_EXE = os.path.join(ROOT, "tests", "cpp", "callers_test")


def build_callers_test():
    libdir = os.path.join(ROOT, "paillier_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "callers_test.cpp"),
                    "-o", CALLERS_EXE, "-L", libdir, "-lpaillier_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    return CALLERS_EXE


def test_cpp_host_helpers_of_the_drawing_callers():
    """The host-only helpers behind the C++ mirror's randomness-drawing callers (no GPU): the byte-string product that gives
    n^2 (range of a proof's r, thresholdkey.go:233) against Python ints, and random_below (crypto/rand.Int's rejection
    sampling, utils.go:26-33) staying below its bound and covering it."""
    import random
    exe = build_callers_test()
    rnd = random.Random(12)
    pairs = [(0, 5), (1, 1), (255, 255), (2 ** 64 - 1, 2 ** 64 - 1), (2 ** 2048 - 1, 2 ** 2048 - 1), (2 ** 6144 - 1, 2 ** 6144 - 1)]
    pairs += [(rnd.getrandbits(a), rnd.getrandbits(b)) for a, b in ((8, 4096), (1024, 1024), (2048, 2048), (3072, 3072), (4096, 17))]
    bounds = [1, 2, 3, 255, 256, 257, 2 ** 64, rnd.getrandbits(2048) | 1 << 2047, 5 << 1000]
    lines = [f"mul {a:x} {b:x}" for a, b in pairs] + [f"below {b:x}" for b in bounds] + [f"osrandom {1 << 128:x}"] * 2
    r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout.split("\n")
    assert [int(x, 16) for x in out[:len(pairs)]] == [a * b for a, b in pairs]
    for b, line in zip(bounds, out[len(pairs):]):
        draws = [int(x, 16) for x in line.split()]
        assert len(draws) == 64 and all(0 <= d < b for d in draws)
        if b >= 256:
            assert len(set(draws)) > 32 and max(draws) >= b // 4          # spread over the range, top bits used
        elif b > 1:
            assert len(set(draws)) > 1
    a, b = ([int(x, 16) for x in out[len(pairs) + len(bounds) + i].split()] for i in (0, 1))
    assert len(set(a + b)) == 128 and all(0 <= v < 1 << 128 for v in a + b) and max(a + b) >> 120      # getrandom(2): no repeats, full range
    assert out[len(pairs) + len(bounds) + 2] == "callers ok"


def test_gob_decoders_agree_on_mutated_streams():
    """Differential fuzz of the two wire-format decoders (untrusted input): 6000 mutations (byte flips, truncations, insertions,
    deletions, trailing bytes) of valid Ciphertext streams through paillier::gob::decode and gobwire.decode_ciphertext.  Neither may
    crash; they accept the same streams with the same (C, Level, EncMethod).  One documented divergence: a negative C decodes in
    Python as in Go (NewCiphertextFromBytes does not range-check, paillier.go:374-390) while the C++ mirror's Int is an unsigned
    magnitude and reports "negative ciphertext value"."""
    import random
    from paillier_b200 import gobwire as W
    exe = os.path.join(ROOT, "tests", "cpp", "gob_test")
    if not os.path.exists(exe):
        test_cpp_gob_wire_format_matches_python_mirror()
    rnd = random.Random(77)
    bases = [W.encode_ciphertext(rnd.getrandbits(b) | 1, rnd.randrange(2), rnd.randrange(3), struct_id=rnd.choice([65, 66, 70, 127, 128, 300]))
             for b in (8, 64, 200, 1024, 4096)]

    def mutate(s):
        s = bytearray(s)
        k = rnd.randrange(6)
        if k == 0:
            for _ in range(rnd.randrange(1, 4)):
                s[rnd.randrange(len(s))] = rnd.randrange(256)
        elif k == 1:
            s = s[:rnd.randrange(len(s))]
        elif k == 2:
            i = rnd.randrange(len(s))
            s[i:i] = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 5)))
        elif k == 3:
            i = rnd.randrange(len(s))
            del s[i:i + rnd.randrange(1, 5)]
        elif k == 4:
            s[rnd.randrange(len(s))] ^= 1 << rnd.randrange(8)
        else:
            s = s + bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 9)))
        return bytes(s)

    cases = [c for c in (mutate(rnd.choice(bases)) for _ in range(6000)) if c]
    r = subprocess.run([exe], input="\n".join("dec " + c.hex() for c in cases) + "\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout.split("\n")
    assert len(out) >= len(cases)
    both = 0
    for c, o in zip(cases, out):
        try:
            v = W.decode_ciphertext(c)
            py = f"{hex(v[0])} {v[1]} {v[2]}"
        except ValueError:                          # every rejection of the Python decoder is a ValueError
            py = None
        if py is None:
            assert o.startswith("error "), (c.hex(), o)
        elif v[0] < 0:
            assert o == "error negative ciphertext value", (c.hex(), o)
        else:
            assert o == py, (c.hex(), py, o)
            both += 1
    assert both > 1000                               # a third of the mutations leave a valid stream: the test is not vacuous
