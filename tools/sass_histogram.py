"""SASS opcode histogram of the hot kernels of libpaillier_b200.so:
    python tools/sass_histogram.py [out.json]
Per kernel: static instruction count, registers, and the counts of the opcodes that say which pipe the kernel lives on
(IMAD.WIDE.U32[.X], DFMA, DADD, IADD3[.X], LOP3, SHFL, LDG/STG, LDS/STS, LDL/STL = spills)."""
import collections, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "paillier_b200", "libpaillier_b200.so")
WANT = re.compile(r"powm_vm|strong_kernel|prod_reduce|crt_combine|modinv_kernel|sha256")
GROUPS = ["IMAD.WIDE.U32.X", "IMAD.WIDE.U32", "IMAD.WIDE", "IMAD", "DFMA", "DADD", "DMUL", "IADD3.X", "IADD3", "LOP3", "SHF", "SHFL", "LEA", "SEL", "ISETP",
          "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "BRA"]


def group(op):
    for g in GROUPS:
        if op == g or op.startswith(g + "."):
            return g
    return "other"


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res_usage = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    regs = {}
    cur = None
    for line in res_usage.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    kernels = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1) if WANT.search(m.group(1)) else None
            if cur:
                kernels[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                kernels[cur][group(m.group(1))] += 1
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    doc = {"library": "paillier_b200/libpaillier_b200.so (cuobjdump -sass, sm_100a)", "kernels": {}}
    for name, hist in sorted(kernels.items()):
        total = sum(hist.values())
        r = regs.get(name, (None, None))
        doc["kernels"][demangle(name)] = {"instructions": total, "registers": r[0], "stack_bytes": r[1],
                                          "opcodes": dict(sorted(hist.items(), key=lambda kv: -kv[1]))}
    text = json.dumps(doc, indent=1)
    if out:
        open(out, "w").write(text + "\n")
        print("wrote", out, len(doc["kernels"]), "kernels")
    else:
        print(text)


if __name__ == "__main__":
    main()
