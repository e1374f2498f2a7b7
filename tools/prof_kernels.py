"""One small launch of each hot kernel family, for ncu / compute-sanitizer captures:
    python tools/prof_kernels.py enc|dec|pdec3072|level2|ddleq|safeprime|zkp|light [count]
Each mode checks its result (round trip or proof verification), so a capture of a wrong kernel cannot pass unnoticed."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paillier_b200 import synth
from paillier_b200.api import ENC_LEVEL_TWO, PublicKey, SecretKey, from_records, to_records

what = sys.argv[1] if len(sys.argv) > 1 else "enc"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 0

if what in ("enc", "dec", "level2", "ddleq", "light"):
    p, q = synth.load_key("paillier_2048")
    sk = SecretKey(p * q, p=p, q=q)
    n = sk.N
    if what in ("enc", "dec"):
        count = count or sk.kernel_shape(1)["resident_groups"]
        m = synth.plaintexts(count, n, sk.w_n); r = synth.randomness(count, n, sk.w_n)
        c = PublicKey.encrypt_with_r_records(sk, m, r)          # the public-key path: r^n mod n^2 in one launch
        assert np.array_equal(sk.decrypt_records(c), m)
    elif what == "light":
        count = count or 65536
        m = synth.plaintexts(count, n, sk.w_n); r = synth.randomness(count, n, sk.w_n)
        c = sk.encrypt_with_r_records(m, r)                     # the key holder's path: r^n over p^2, q^2
        s = sk.modmul_records(1, c, c[::-1].copy(), sk.w_n2)
        assert s.size == c.size
    elif what == "level2":
        count = count or 4096
        w2 = sk.w_n2
        ms = from_records(synth.random_records(count, w2, (n * n).bit_length() - 1, stream=51), w2)
        rs = from_records(synth.randomness(count, n, sk.w_n), sk.w_n)
        cts = sk.EncryptWithRAtLevelBatch(ms, rs, ENC_LEVEL_TWO)
        assert sk.DecryptBatch(cts) == ms
    else:
        count = count or 64
        secpar = 8
        ints = lambda a: from_records(a, sk.w_n)
        rr = ints(synth.randomness(count * (4 + 2 * secpar), n, sk.w_n, synth.SEED + 7))
        inner = sk.EncryptWithRBatch(ints(synth.plaintexts(count, n, sk.w_n)), rr[:count])
        ct1 = sk.EncryptWithRAtLevelBatch([c.C for c in inner], rr[count:2 * count], ENC_LEVEL_TWO)
        As, Bs = rr[2 * count:3 * count], rr[3 * count:4 * count]
        ct2 = sk.NestedRandomizeWithBatch(ct1, As, Bs)
        xs = [rr[4 * count + i * secpar:4 * count + (i + 1) * secpar] for i in range(count)]
        ys = [rr[(4 + secpar) * count + i * secpar:(4 + secpar) * count + (i + 1) * secpar] for i in range(count)]
        proofs = sk.ProveDDLEQBatch(secpar, ct1, ct2, As, Bs, xs, ys)
        assert all(sk.VerifyDDLEQProofBatch(ct1, ct2, proofs))
elif what in ("pdec3072", "zkp"):
    from paillier_b200.keygen import ThresholdKeyGenerator
    bits = 3072
    p, q = synth.load_key(f"threshold_{bits}")
    n = p * q
    keys = ThresholdKeyGenerator(bits, 8, 5, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys()
    tsk = keys[0]
    count = count or tsk.kernel_shape(1)["resident_groups"]
    c = tsk.encrypt_with_r_records(synth.plaintexts(count, n, tsk.w_n), synth.randomness(count, n, tsk.w_n))
    if what == "pdec3072":
        decs = [k.partial_decrypt_records(c) for k in keys[:5]]
        got = tsk.combine_records([k.ID for k in keys[:5]], np.concatenate(decs))
        assert np.array_equal(got, synth.plaintexts(count, n, tsk.w_n))
    else:
        r = synth.random_records(count, tsk.w_n2, (n * n).bit_length() - 1, stream=5)
        dec, e, z = tsk.zkp_prove_records(c, r)
        assert tsk.verify_proof_records(tsk.ID, c, dec, e, z).all()
elif what == "safeprime":
    from paillier_b200.keygen import safe_prime_scan
    count = count or 16384
    raw = synth.random_records(count, 128, 1024, stream=41).tobytes()
    ps, qs, ok = safe_prime_scan(1024, raw)
    assert len(ok) == count
else:
    raise SystemExit("unknown mode " + what)
print("ok", what, count)
