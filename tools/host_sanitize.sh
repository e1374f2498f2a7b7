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
