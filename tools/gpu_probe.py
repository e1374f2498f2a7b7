"""Shape sweep on the GPU box: kernel-only time of EncryptWithR / Decrypt / PartialDecrypt per (TPI, L)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from paillier_b200 import synth
from paillier_b200._lib import lib, check
from paillier_b200.api import SecretKey


def kernel_ms(key):
    ms = C.c_float()
    check(lib.pgpu_ctx_last_kernel_ms(key._ctx, C.byref(ms)), key._ctx)
    return ms.value


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
    p, q = synth.load_key("paillier_2048")
    n = p * q
    res = []
    for s128 in ("8,16", "16,8", "4,32", "32,4"):
        for s64 in ("4,16", "8,8", "2,16"):
            if s128 != "8,16" and s64 != "4,16":
                continue
            os.environ["PGPU_SHAPE_128"] = s128
            os.environ["PGPU_SHAPE_64"] = s64
            sk = SecretKey(n, p=p, q=q)
            check(lib.pgpu_ctx_enable_timing(sk._ctx, 1), sk._ctx)
            m = synth.plaintexts(count, n, sk.w_n)
            r = synth.randomness(count, n, sk.w_n)
            c = sk.encrypt_with_r_records(m, r)       # warm-up
            c = sk.encrypt_with_r_records(m, r)
            t_enc = kernel_ms(sk)
            d = sk.decrypt_records(c)
            d = sk.decrypt_records(c)
            t_dec = kernel_ms(sk)
            assert np.array_equal(d, m)
            S, sq, mu = sk.program_cost(0)
            macs = sq * (1.5 * S * S + 1.5 * S) + mu * (2 * S * S + S)
            row = {"shape128": s128, "shape64": s64, "count": count, "enc_ms": t_enc, "dec_ms": t_dec,
                   "enc_per_s": count / t_enc * 1e3, "dec_per_s": count / t_dec * 1e3,
                   "enc_prog": [S, sq, mu], "enc_tmacs": macs * count / t_enc * 1e3 / 1e12}
            print(json.dumps(row), flush=True)
            res.append(row)
            sk.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
