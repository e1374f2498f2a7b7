#!/bin/bash
# Host-side memory / undefined-behaviour check (no GPU needed): builds the library's HOST code with AddressSanitizer and
# UndefinedBehaviorSanitizer in a scratch copy and runs the CPU tests that execute it -- the key set-up big integers
# (bn_host.hpp through pgpu_selftest_bn), the exponentiation-program compiler (pgpu_selftest_program), every export with
# null / zero arguments -- plus the C++ gob decoder on 6000 mutated streams.  compute-sanitizer (device code) is closed on
# the GPU pool (profiles/r02_sanitizer_unavailable.txt); this covers the other half of the library.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/pgpu_asan}
rm -rf "$W" && mkdir -p "$W/paillier_b200" "$W/include"
cp -r "$ROOT/paillier_b200/csrc" "$W/paillier_b200/" && cp "$ROOT"/include/* "$W/include/" && rm -f "$W"/paillier_b200/csrc/*.o
SAN="-Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined"
sed -i "s/NVFLAGS := \$(ARCH) -O3/NVFLAGS := \$(ARCH) -O1 $SAN -Xcompiler -fno-omit-frame-pointer/; s/-cudart shared -ldl/-cudart shared -ldl $SAN/" "$W/paillier_b200/csrc/Makefile"
make -C "$W/paillier_b200/csrc" -j8 > "$W/build.log" 2>&1
export ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
cd "$ROOT"
PGPU_LIB_PATH="$W/paillier_b200/libpaillier_b200.so" python -m pytest tests/test_abi_and_host.py tests/test_host_programs.py -q \
    -k "not header_symbols and not cpp_mirror and not go_binding"
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -I include tests/cpp/gob_test.cpp -o "$W/gob_test_asan" \
    -L paillier_b200 -lpaillier_b200 -Wl,-rpath,"$ROOT/paillier_b200"
python - "$W/gob_test_asan" <<'PY'
import random, subprocess, sys
sys.path.insert(0, ".")
from paillier_b200 import gobwire as W
rnd = random.Random(21)
bases = [W.encode_ciphertext(rnd.getrandbits(b) | 1, rnd.randrange(2), rnd.randrange(3)) for b in (8, 64, 1024, 4096)]
cases = []
for _ in range(6000):
    s = bytearray(rnd.choice(bases))
    for _ in range(rnd.randrange(1, 5)):
        k = rnd.randrange(3)
        if k == 0: s[rnd.randrange(len(s))] = rnd.randrange(256)
        elif k == 1 and len(s) > 1: s = s[:rnd.randrange(1, len(s))]
        else: i = rnd.randrange(len(s)); s[i:i] = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 9)))
    cases.append(bytes(s))
r = subprocess.run([sys.argv[1]], input="\n".join("dec " + c.hex() for c in cases) + "\n", capture_output=True, text=True)
print("gob decoder under ASan/UBSan: rc", r.returncode, "stderr bytes", len(r.stderr), "lines", len(r.stdout.split("\n")) - 1)
sys.exit(r.returncode or (1 if r.stderr else 0))
PY

# ---- the multiplier sources under ThreadSanitizer: one thread per lane of an emulated warp (tests/cpp/cuda_host_shim.h), so a
# missing __syncwarp() in the dedicated squaring's shared-memory exchange (csrc/mont.cuh: Mont::sqr) is a real data race
g++ -std=c++20 -O1 -g -frounding-math -pthread -fsanitize=thread -Wno-unknown-pragmas tests/cpp/mont_host_test.cpp -o "$W/mont_host_tsan"
unset LD_PRELOAD
python - "$W/mont_host_tsan" "$W" <<'PY'
import random, re, subprocess, sys
exe, W = sys.argv[1], sys.argv[2]
rnd = random.Random(7)
src = open("paillier_b200/csrc/powm.cu").read()
s32 = re.findall(r"X\((\d+),\s*(\d+)\)", src[src.index("#define PGPU_FOR_EACH_SHAPE(X)"):src.index("// FP64-pipe shapes")])
s52 = re.findall(r"X\((\d+),\s*(\d+),\s*(\d+)\)", src[src.index("#define PGPU_FOR_EACH_SHAPE52(X)"):src.index("cudaError_t vm_launch")])
lines = []
for t, l in s32:
    t, l = int(t), int(l)
    bits, g = 32 * t * l, 32 // t
    n = rnd.getrandbits(bits) | 1 | 1 << (bits - 1)
    lines.append(f"m32 {t} {l} {int(t == 4 and l % 8 == 0 and l <= 16)} {n:x} {g} " + " ".join(f"{rnd.randrange(n):x} {rnd.randrange(n):x}" for _ in range(g)))
for t, l, s in s52:
    t, l, s = int(t), int(l), int(s)
    n = rnd.getrandbits(32 * s) | 1 | 1 << (32 * s - 1)
    lines.append(f"m52 {t} {l} {s} {n:x} {32 // t} " + " ".join(f"{rnd.randrange(n):x} {rnd.randrange(n):x}" for _ in range(32 // t)))
inp = "\n".join(lines) + "\n"
r = subprocess.run([exe], input=inp, capture_output=True, text=True, env={"TSAN_OPTIONS": "halt_on_error=0"})
races = r.stderr.count("WARNING: ThreadSanitizer")
print(f"multiplier sources under ThreadSanitizer ({len(s32)} integer + {len(s52)} FP64 shapes, mul/sqr/add/sub on every lane group): rc {r.returncode}, {races} reports")
# negative control: the same harness against a copy of mont.cuh with ONE __syncwarp() of the squaring removed must be reported
import os, shutil
neg = os.path.join(W, "neg")
shutil.rmtree(neg, ignore_errors=True)
os.makedirs(os.path.join(neg, "paillier_b200", "csrc")); os.makedirs(os.path.join(neg, "tests", "cpp"))
for f in ("mont.cuh", "mont52.cuh"): shutil.copy(os.path.join("paillier_b200", "csrc", f), os.path.join(neg, "paillier_b200", "csrc", f))
for f in ("cuda_host_shim.h", "mont_host_test.cpp"): shutil.copy(os.path.join("tests", "cpp", f), os.path.join(neg, "tests", "cpp", f))
p = os.path.join(neg, "paillier_b200", "csrc", "mont.cuh")
s = open(p).read()
old = "block_product_to_smem<L>(a, sa, round == 0 ? gb + ((t + 1) & 3) : wl, 0);\n                __syncwarp();\n"
assert old in s
open(p, "w").write(s.replace(old, old.replace("                __syncwarp();\n", ""), 1))
subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-frounding-math", "-pthread", "-fsanitize=thread", "-Wno-unknown-pragmas",
                os.path.join(neg, "tests", "cpp", "mont_host_test.cpp"), "-o", os.path.join(W, "mont_host_tsan_neg")], check=True)
r2 = subprocess.run([os.path.join(W, "mont_host_tsan_neg")], input=inp, capture_output=True, text=True, env={"TSAN_OPTIONS": "halt_on_error=0"})
neg_races = r2.stderr.count("WARNING: ThreadSanitizer")
first = re.search(r"Read of size.*?\n\s+#0 (.*?) \(", r2.stderr, re.S)
print(f"negative control (one __syncwarp of Mont::sqr removed): rc {r2.returncode}, {neg_races} reports; first: {first.group(1)[:160] if first else None}")
sys.exit(1 if (r.returncode or races or not neg_races) else 0)
PY
# the whole emulation test file under ThreadSanitizer: the multipliers as above plus the INTERPRETER's source (csrc/vm_run.cuh)
# running library-compiled programs -- table, dump and output accesses of the resident groups
PGPU_EMU_CXXFLAGS="-fsanitize=thread -g" TSAN_OPTIONS="halt_on_error=0" python -m pytest tests/test_mont_host_emulation.py -q
g++ -std=c++20 -O1 -frounding-math -pthread -Wno-unknown-pragmas tests/cpp/mont_host_test.cpp -o tests/cpp/mont_host_test   # leave the plain build behind
