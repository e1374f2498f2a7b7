#!/usr/bin/env python
"""bench.py -- BASELINE.json configs[1]: 2048-bit n, a batch of 2^20 EncryptWithR (fixed, seeded r for the
bit-exact check) followed by CRT Decrypt of the same batch, on N B200s (one process per GPU, the batch
sharded across ranks with no collective on the data path: weak scaling, 2^20 items per GPU).

    python bench.py --gpus N --steps K --warmup W [--count C] [--impl reference]

One "step" = one pass of the hot path over one batch: EncryptWithR over C items, then Decrypt over the C
ciphertexts.  `value` = items through the whole step per second, summed over ranks, inputs resident in HBM.
`e2e` = the same step through the host-buffer C-ABI calls (pgpu_encrypt_with_r / pgpu_decrypt) with pinned
host buffers, H2D and D2H copies inside the timed region.  `breakdown` carries the headline enc/s, dec/s and
partial-dec/s separately; `strong` is one 2^20 batch split over the N ranks (configs[1]'s "then sharded across 2/4/8");
`config4` is BASELINE configs[3] (3072-bit threshold key, 8 shares / threshold 5, PartialDecrypt + proofs for every
ciphertext, NCCL all-gather, proof verification and Combine of each rank's slice) with device time per phase;
`cpu_baselines` times the libgmp call sequences of the other rows on the host cores.  `--impl reference` times the libgmp
restatement of the reference's call sequence (oracle/gmp_ref.c: a stand-in for the Go package, which cannot be built in
this image) on all host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "2048-bit Paillier EncryptWithR + CRT Decrypt items/s"
UNIT = "items/s"
WORKLOAD = "config[1]: 2048-bit n, EncryptWithR (seeded r) + CRT Decrypt over a batch, sharded across GPUs"


def fp64_peak() -> dict | None:
    """Measured DFMA issue rate of this GPU (tools/dfma_peak.cu)"""
    exe = os.path.join(ROOT, "tools", "dfma_peak")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120, check=True).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception:
        return None


def mont_macs(S: int, n_sqr: int, n_mul: int) -> float:
    """SURVEY.md 8(d) accounting: Montgomery mul = 2s^2+s, sqr = 1.5s^2+1.5s MAC32 (as if squarings were dedicated)."""
    return n_sqr * (1.5 * S * S + 1.5 * S) + n_mul * (2.0 * S * S + S)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        mhz = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": sorted(reasons)}


def run_reference(args) -> None:
    """The reference's own CPU implementation of the path (libgmp call sequence of paillier.go:206-218 and :292-303),
    all host cores, a bounded sample of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import gmp_ref as G
    from paillier_b200 import synth
    p, q = synth.load_key("paillier_2048")
    n, lam = p * q, (p - 1) * (q - 1)
    w_n = 256
    cores = G.cores()
    sample = args.ref_sample or 64 * cores
    m = synth.plaintexts(sample, n, w_n)
    r = synth.randomness(sample, n, w_n)

    def step():
        c = G.encrypt_with_r(n, m, r, w_n, threads=cores)
        d = G.decrypt(n, lam, c, w_n, threads=cores)
        return c, d

    for _ in range(max(args.warmup, 1)):
        c, d = step()
    assert np.array_equal(d, m)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    t1 = time.perf_counter(); G.encrypt_with_r(n, m, r, w_n, threads=cores); te = time.perf_counter() - t1
    t1 = time.perf_counter(); G.decrypt(n, lam, c, w_n, threads=cores); td = time.perf_counter() - t1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (libgmp u64)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "key": "tests/golden/keys.json:paillier_2048", "items_per_step": sample,
                   "note": "a bounded sample per step (the B200 arm runs 2^20 items per GPU per step): items/s is size-independent on the CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} items/step: libgmp mpz_powm call sequence of paillier.go:213-216 and :296-300 "
                                   "(no g=n+1 shortcut, no CRT); stand-in for the Go package (no Go toolchain in the image)"},
        "breakdown": {"enc_per_s": sample / te, "dec_per_s": sample / td},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def imad_peak() -> dict | None:
    """Measured IMAD.WIDE issue rate of this GPU (tools/imad_peak.cu), the roofline denominator."""
    exe = os.path.join(ROOT, "tools", "imad_peak")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120, check=True).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception:
        return None


def config4_leg(args, dist, rank, world, local, dev, barrier, max_over_ranks):
    """BASELINE configs[3]: 3072-bit n threshold key, 8 shares / threshold 5.  The 8 share-holders are spread over the N
    GPUs (8/N per GPU; one per GPU at N = 8).  Every share-holder computes PartialDecrypt (thresholdkey.go:192-201) and the
    proof (thresholdkey.go:225-255) for ALL ciphertexts, one NCCL all-gather per field lands [share][ciphertext], every rank
    verifies the 8 proofs of its ciphertext slice (:278-311) and combines it (:149-172).  Timed on the device, max over ranks."""
    import random
    import numpy as np
    import torch
    from paillier_b200 import synth
    from paillier_b200.keygen import ThresholdKeyGenerator
    from paillier_b200.multi import gpu_threshold_round_shares, shard_range

    l, w = 8, 5
    k = l // world
    tp3, tq3 = synth.load_key("threshold_3072")
    n3 = tp3 * tq3
    keys = ThresholdKeyGenerator(3072, l, w, rng=random.Random(synth.SEED)).with_safe_primes(tp3, tq3).GenerateKeys(device=local)
    mine = keys[rank * k:(rank + 1) * k]
    for t in keys:
        if t not in mine:
            t.close()
    t0 = mine[0]
    count = args.c4_count or min(1 << 18, (1 << 15) * world)
    w_n, w2 = t0.w_n, t0.w_n2
    # the same seeded ciphertext batch on every rank (valid ciphertexts: EncryptWithR of seeded plaintexts)
    m = synth.plaintexts(count, n3, w_n, synth.SEED + 4)
    c_dev = torch.from_numpy(t0.encrypt_with_r_records(m, synth.randomness(count, n3, w_n, synth.SEED + 4))).to(dev)
    zr = [torch.from_numpy(synth.random_records(count, w2, 2 * n3.bit_length() - 2, synth.SEED + 4, stream=70 + t.ID)).to(dev) for t in mine]
    d = dist if world > 1 else None
    small = min(count, 256)
    gpu_threshold_round_shares(d, mine, c_dev[:small * w2], small, world, rank, [z[:small * w2] for z in zr])        # warm-up: programs, tables
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    keep = {}
    plain, (lo, hi), phases = gpu_threshold_round_shares(d, mine, c_dev, count, world, rank, zr, keep=keep)
    e1.record(torch.cuda.current_stream(dev))
    barrier()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    ok = bool(np.array_equal(plain.cpu().numpy(), m[lo * w_n:hi * w_n]))
    bad = max_over_ranks(0.0 if ok else 1.0)
    ph = {name: max_over_ranks(v) for name, v in phases.items()}
    out = None
    if rank == 0:
        # parity of a slice of share-holder 1's output against the libgmp call sequence (thresholdkey.go:192-255)
        from oracle import gmp_ref as G
        ns = min(count, 32)
        cs = c_dev[:ns * w2].cpu().numpy()
        rs = zr[0][:ns * w2].cpu().numpy()
        rd, re_, rz = G.pdec_zkp(n3, t0.Share, l, t0.VerificationKey, cs, rs, w2, t0.w_z)
        par = (np.array_equal(rd, keep["dec"][:ns * w2].cpu().numpy()) and np.array_equal(re_, keep["e"][:ns * 32].cpu().numpy())
               and np.array_equal(rz, keep["z"][:ns * t0.w_z].cpu().numpy()))
        S3, q3, m3 = t0.program_cost(2)
        out = {
            "workload": "config[3]: 3072-bit n threshold key, 8 shares (threshold 5) on %d GPU(s), %d per GPU; PartialDecrypt + proof for "
                        "every ciphertext, all-gather, VerifyProof x 8 and Combine of each rank's slice" % (world, k),
            "ciphertexts": count, "ciphertexts_note": "2^18 per share-holder at 8 GPUs; min(2^18, 2^15 * N) otherwise so that the leg fits the bench budget",
            "ms": total_ms, "ciphertexts_per_s": count / (total_ms * 1e-3), "partial_decryptions_per_s": l * count / (total_ms * 1e-3),
            "phases_ms": ph, "phases_note": "pdec is empty: with proofs the partial decryption comes out of the proof's launch (c^(2*delta*s) and (c^4)^r "
                                            "share their squarings); verify raises c^4 to the 8 share-holders' Z with shared squarings",
            "all_gather_share_of_step": ph["all_gather"] / total_ms if total_ms else None,
            "all_gather_bytes_per_gpu": {"sent": k * count * (w2 + 32 + t0.w_z), "received": l * count * (w2 + 32 + t0.w_z)},
            "all_proofs_verified_on_rank0": bool(keep["ok"].all().item()) if keep.get("ok") is not None else None,
            "all_plaintexts_recovered": bad == 0.0,
            "oracle_parity": {"items": ns, "fields": "c_i, E, Z of share-holder 1 against oracle/gmp_ref.c", "equal": bool(par)},
            "pdec_program": {"limbs": S3, "sqr": q3, "mul": m3, "mac32_per_item": mont_macs(S3, q3, m3)},
            "kernel": t0.kernel_shape(1),
        }
        assert par, "config 4: partial decryption / proof differs from the libgmp oracle"
    assert bad == 0.0, "config 4: recovered plaintexts differ from the inputs"
    for t in mine:
        t.close()
    return out


def cpu_baselines_leg(sk, n, p, q, w_n, w_n2, c_host_np, extras_keep, breakdown):
    """libgmp call sequences of the other rows (oracle/gmp_ref.c), all host cores, bounded samples; each entry doubles as a
    bit-exact check of the GPU result where the GPU result is at hand."""
    import random
    import numpy as np
    from oracle import gmp_ref as G
    from paillier_b200 import synth
    cores = G.cores()
    out = {"cores": cores, "kind": "port", "note": "libgmp call sequence of the reference per row, one thread per core; samples sized for 1-4 s each"}
    # PartialDecrypt 2048 / 3072 (thresholdkey.go:192-201)
    for bits, ns in ((2048, 8 * cores), (3072, 4 * cores)):
        tp, tq = synth.load_key(f"threshold_{bits}")
        nn = tp * tq
        wn2 = {2048: 512, 3072: 768}[bits]
        share = random.Random(bits).randrange(nn * ((tp - 1) // 2) * ((tq - 1) // 2))
        cs = synth.random_records(ns, wn2, 2 * nn.bit_length() - 2, stream=81)
        t0 = time.perf_counter()
        G.partial_decrypt(nn, share, 8, cs, wn2, threads=cores)
        out[f"pdec{bits}_per_s"] = ns / (time.perf_counter() - t0)
        # ZKP prove + verify at the same size (thresholdkey.go:225-311)
        v = pow(random.Random(bits + 1).randrange(2, nn * nn), 2, nn * nn)
        vi = pow(v, 40320 * share, nn * nn)
        nz = max(cores, ns // 4)
        rs = synth.random_records(nz, wn2, 2 * nn.bit_length() - 2, stream=82)
        wz = wn2 + 64
        t0 = time.perf_counter()
        dec, e, z = G.pdec_zkp(nn, share, 8, v, cs[:nz * wn2], rs, wn2, wz, threads=cores)
        t1 = time.perf_counter()
        okv = G.zkp_verify(nn, v, vi, cs[:nz * wn2], dec, e, z, wn2, wz, threads=cores)
        t2 = time.perf_counter()
        assert bool(okv.all()), "libgmp oracle: ZKP verification rejected its own proof"
        out[f"pdec_zkp_prove{bits}_per_s"] = nz / (t1 - t0)
        out[f"pdec_zkp_verify{bits}_per_s"] = nz / (t2 - t1)
    # encrypted dot product with 64-bit scalars (operations.go:11-64): ConstMult per term + Add fold
    nd = 256 * cores
    ks = synth.scalars_u64(nd)
    t0 = time.perf_counter()
    ref_dot = G.dot_u64(n * n, c_host_np[:nd * w_n2], w_n2, ks, threads=cores)
    out["dot_u64_terms_per_s"] = nd / (time.perf_counter() - t0)
    out["dot_u64_sample"] = nd
    # DDLEQ verification (ddleq.go:129-153) of the GPU's own proofs: baseline and parity in one
    if "ddleq" in extras_keep:
        dn, dsecpar, c1r, c2r, xr, yr, al_r, e_r, f_r = extras_keep["ddleq"]
        ns = max(1, min(dn, 2 * cores // dsecpar or 1))
        w3 = sk.w_n3
        t0 = time.perf_counter()
        okr = G.ddleq_verify(n, dsecpar, c1r[:ns * w3], c2r[:ns * w3], xr[:ns * dsecpar * w_n], yr[:ns * dsecpar * w_n], al_r[:ns * dsecpar * w3],
                             e_r[:ns * dsecpar * w_n2], f_r[:ns * dsecpar * w3], w_n, w_n2, w3, threads=cores)
        out["ddleq_verify_instances_per_s"] = ns * dsecpar / (time.perf_counter() - t0)
        assert bool(okr.all()), "libgmp oracle rejected a DDLEQ proof made on the GPU"
    # safe-prime candidates (safe_prime.go:170-263): same byte strings as the GPU leg, decisions compared
    if "safe_prime" in extras_keep:
        rawb, gpu_ok = extras_keep["safe_prime"]
        nbc = (1024 - 1 + 7) // 8
        ns = 2048 * cores
        t0 = time.perf_counter()
        okc = G.safe_prime_scan(1024, rawb[:ns * nbc].tobytes(), threads=cores)
        out["safe_prime_candidates_per_s"] = ns / (time.perf_counter() - t0)
        assert np.array_equal(okc, gpu_ok[:ns]), "safe-prime decisions differ between the GPU and the libgmp oracle"
    return out, ref_dot



def config1_leg(local):
    """BASELINE configs[0]: 1024-bit n, KeyGen + Encrypt / Decrypt / Add round trip over 10^4 random plaintexts ON THE CPU
    (paillier_test.go:52-63, operations_test.go:11-28).  Two CPU lines as BASELINE.md promised -- the libgmp call sequence on
    all cores and the Python restatement on one core (a sample) -- and the same batch through the C ABI beside them."""
    import random
    import numpy as np
    from oracle import gmp_ref as G
    from oracle import paillier_ref as R
    from paillier_b200 import synth
    from paillier_b200.api import PublicKey, SecretKey, from_records
    count = 10_000
    p, q = synth.load_key("paillier_1024")
    n, lam = p * q, (p - 1) * (q - 1)
    w = 128
    cores = G.cores()
    m = synth.plaintexts(count, n, w, synth.SEED + 11)
    r = synth.randomness(count, n, w, synth.SEED + 11)
    t0 = time.perf_counter(); c = G.encrypt_with_r(n, m, r, w, threads=cores)
    t1 = time.perf_counter(); d = G.decrypt(n, lam, c, w, threads=cores)
    t2 = time.perf_counter(); tot = G.add_reduce(n * n, c, 2 * w, threads=cores)
    t3 = time.perf_counter()
    assert np.array_equal(d, m)
    out = {"workload": "config[0]: 1024-bit n, 10^4 plaintexts, Encrypt + Decrypt + Add round trip", "items": count,
           "libgmp": {"cores": cores, "enc_per_s": count / (t1 - t0), "dec_per_s": count / (t2 - t1), "add_terms_per_s": count / (t3 - t2),
                      "round_trip_items_per_s": count / (t3 - t0)}}
    # Python restatement (CPython ints), one core, first 200 items
    ns = 200
    osk, opk = R.keygen_from_primes(p, q)
    ms, rs = from_records(m[:ns * w], w), from_records(r[:ns * w], w)
    t0 = time.perf_counter(); ocs = [R.encrypt_with_r(opk, a, b) for a, b in zip(ms, rs)]
    t1 = time.perf_counter(); ods = [R.decrypt(osk, ct) for ct in ocs]
    t2 = time.perf_counter(); osum = R.add(opk, *ocs)
    t3 = time.perf_counter()
    assert ods == ms and [ct.C for ct in ocs] == from_records(c[:ns * 2 * w], 2 * w)
    out["python_oracle"] = {"cores": 1, "sample": ns, "enc_per_s": ns / (t1 - t0), "dec_per_s": ns / (t2 - t1), "add_terms_per_s": ns / (t3 - t2),
                            "round_trip_items_per_s": ns / (t3 - t0)}
    # KeyGen (paillier.go:106-179) restated on the CPU: two 512-bit primes = 3 mod 4 by trial + Miller-Rabin (oracle's ProbablyPrime stand-in)
    rng = random.Random(synth.SEED + 12)
    t0 = time.perf_counter()
    primes = []
    while len(primes) < 2:
        cand = rng.getrandbits(512) | (3 << 510) | 3
        if R._is_probable_prime(cand) and cand not in primes:
            primes.append(cand)
    out["python_oracle"]["keygen_s"] = time.perf_counter() - t0
    # the same batch on the GPU through the host-buffer C ABI
    sk = SecretKey(n, p=p, q=q, device=local)
    warm = PublicKey.encrypt_with_r_records(sk, m[:64 * w], r[:64 * w])          # programs, tables, staging buffers
    sk.decrypt_records(warm); sk.add_reduce_records(warm)
    t0 = time.perf_counter(); gc = PublicKey.encrypt_with_r_records(sk, m, r)
    t1 = time.perf_counter(); gd = sk.decrypt_records(gc)
    t2 = time.perf_counter(); gt = sk.add_reduce_records(gc)
    t3 = time.perf_counter()
    assert np.array_equal(gc, c) and np.array_equal(gd, m) and np.array_equal(gt, tot), "config 1: GPU results differ from libgmp"
    assert R.decrypt(osk, R.Ciphertext(from_records(gt, 2 * w)[0])) == sum(from_records(m, w)) % n
    out["b200_e2e"] = {"enc_per_s": count / (t1 - t0), "dec_per_s": count / (t2 - t1), "add_terms_per_s": count / (t3 - t2),
                       "round_trip_items_per_s": count / (t3 - t0), "bit_exact_with_libgmp": True}
    sk.close()
    return out


def small_batch_leg(sk, m_np, r_np, n, lam, w_n, w_n2):
    """Latency of ONE blocking host-buffer call at small batch sizes -- what a caller of the reference's scalar API sees when
    it switches to the batch call with few items (paillier.go:185-187,292-303 are one item per call) -- beside one libgmp
    thread doing the same items one after the other.  Pageable host memory (numpy), best of 5 calls after a warm-up call."""
    import numpy as np
    from oracle import gmp_ref as G
    from paillier_b200._lib import check, lib
    from paillier_b200.api import PublicKey
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    rows = {}
    for cnt in (1, 16, 256, 4096):
        m = m_np[:cnt * w_n].copy(); r = r_np[:cnt * w_n].copy()            # copies: pageable memory, like a Go slice
        c = np.empty(cnt * w_n2, dtype=np.uint8); d = np.empty(cnt * w_n, dtype=np.uint8)
        be = bd = 1e9
        for rep in range(6):
            t0 = time.perf_counter(); check(lib.pgpu_encrypt_with_r(sk._ctx, cnt, ptr(m), ptr(r), ptr(c)), sk._ctx)
            t1 = time.perf_counter(); check(lib.pgpu_decrypt(sk._ctx, cnt, ptr(c), ptr(d)), sk._ctx)
            t2 = time.perf_counter()
            if rep:
                be, bd = min(be, t1 - t0), min(bd, t2 - t1)
        assert np.array_equal(d, m), "small batch: Decrypt(Encrypt(m)) != m"
        row = {"encrypt_ms_per_call": be * 1e3, "decrypt_ms_per_call": bd * 1e3, "encrypt_items_per_s": cnt / be, "decrypt_items_per_s": cnt / bd}
        if cnt <= 256:
            k = min(cnt, 16)
            t0 = time.perf_counter(); cref = G.encrypt_with_r(n, m[:k * w_n], r[:k * w_n], w_n, threads=1)
            t1 = time.perf_counter(); G.decrypt(n, lam, cref, w_n, threads=1)
            t2 = time.perf_counter()
            assert np.array_equal(cref, c[:k * w_n2]), "small batch: GPU ciphertexts differ from the libgmp oracle"
            row["libgmp_1_thread_encrypt_ms_per_item"] = (t1 - t0) / k * 1e3
            row["libgmp_1_thread_decrypt_ms_per_item"] = (t2 - t1) / k * 1e3
        rows[str(cnt)] = row
    return {"workload": "one blocking pgpu_encrypt_with_r / pgpu_decrypt call per batch, 2048-bit n, pageable host buffers", "batch": rows,
            "note": "a single item occupies one lane group of one SM for a whole exponentiation: the GPU call pays that latency once per "
                    "batch, libgmp pays its per-item time for every item"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only-small-batch", action="store_true", help="run the small-batch latency leg alone and print it")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--count", type=int, default=1 << 20, help="items per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="items of the cpu_baseline sample (default 48 per core)")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu_baseline / partial-decrypt extras")
    ap.add_argument("--c4-count", type=int, default=0, help="config 4: ciphertexts per share-holder (default min(2^18, 2^15 * N))")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--strong-count", type=int, default=1 << 20, help="items of the strong-scaling batch (split over the ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from paillier_b200 import synth
    from paillier_b200._lib import check, lib
    from paillier_b200.api import SecretKey, ThresholdSecretKey

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up; stdout must carry exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    p, q = synth.load_key("paillier_2048")
    n = p * q
    sk = SecretKey(n, p=p, q=q, device=local)
    # the engine enqueues on this (non-default) torch stream, so torch.cuda.Event timing sees its kernels
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    check(lib.pgpu_ctx_set_stream(sk._ctx, C.c_void_p(stream.cuda_stream)), sk._ctx)
    count, w_n, w_n2 = args.count, sk.w_n, sk.w_n2
    if args.only_small_batch:
        k = 4096
        print(json.dumps(small_batch_leg(sk, synth.plaintexts(k, n, w_n, synth.SEED), synth.randomness(k, n, w_n, synth.SEED), n, (p - 1) * (q - 1),
                                         w_n, w_n2)), flush=True)
        sk.close()
        return

    # seeded synthetic inputs (different stream of the same generator per rank), pinned on the host
    seed = synth.SEED + rank
    m_host = torch.from_numpy(synth.plaintexts(count, n, w_n, seed)).pin_memory()
    r_host = torch.from_numpy(synth.randomness(count, n, w_n, seed)).pin_memory()
    m_dev, r_dev = m_host.to(dev), r_host.to(dev)
    c_dev = torch.empty(count * w_n2, dtype=torch.uint8, device=dev)
    d_dev = torch.empty(count * w_n, dtype=torch.uint8, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())

    def encrypt():
        check(lib.pgpu_encrypt_with_r_dev(sk._ctx, count, vp(m_dev), vp(r_dev), vp(c_dev)), sk._ctx)

    def decrypt():
        check(lib.pgpu_decrypt_dev(sk._ctx, count, vp(c_dev), vp(d_dev)), sk._ctx)

    def step():
        encrypt()
        decrypt()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    assert torch.equal(d_dev, m_dev), "Decrypt(Encrypt(m)) != m"

    l0 = sk.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for k in range(args.steps):
        encrypt()
        ev[2 * k + 1].record(stream)
        decrypt()
        ev[2 * k + 2].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sk.launch_count() - l0
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    enc_ms = max_over_ranks(sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)) / args.steps)
    dec_ms = max_over_ranks(sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(args.steps)) / args.steps)
    ms_per_step = total_ms / args.steps
    value = world * count / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    pdec = None
    if not args.no_extras:
        c_host = torch.empty(count * w_n2, dtype=torch.uint8).pin_memory()
        d_host = torch.empty(count * w_n, dtype=torch.uint8).pin_memory()
        hp = lambda t: C.c_void_p(t.data_ptr())

        def e2e_step():
            check(lib.pgpu_encrypt_with_r(sk._ctx, count, hp(m_host), hp(r_host), hp(c_host)), sk._ctx)
            check(lib.pgpu_decrypt(sk._ctx, count, hp(c_host), hp(d_host)), sk._ctx)

        e2e_step()
        e2e_steps = max(1, min(args.steps, 2))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        assert torch.equal(d_host, m_host), "e2e: Decrypt(Encrypt(m)) != m"
        e2e = {"value": world * count / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": count * (2 * w_n + w_n2), "d2h_bytes_per_step": count * (w_n2 + w_n), "steps": e2e_steps}
        del c_host, d_host

    # ---- roofline of the dominant kernel (powm_vm, the EncryptWithR launch) and the other headline rates
    S, n_sqr, n_mul = sk.program_cost(0)
    enc_macs = mont_macs(S, n_sqr, n_mul)
    Sd, d_sqr, d_mul = sk.program_cost(1)
    dec_macs = mont_macs(Sd, d_sqr, d_mul)
    breakdown = {"enc_per_s": world * count / (enc_ms * 1e-3), "dec_per_s": world * count / (dec_ms * 1e-3),
                 "enc_ms": enc_ms, "dec_ms": dec_ms,
                 "enc_program": {"limbs": S, "sqr": n_sqr, "mul": n_mul, "mac32_per_item": enc_macs},
                 "dec_program": {"limbs": Sd, "sqr": d_sqr, "mul": d_mul, "mac32_per_item": dec_macs}}

    extras_keep = {}
    sk_resident = sk.kernel_shape(1)["resident_groups"]
    if not args.no_extras:
        # EncryptWithR by the key holder (r^n over p^2 and q^2, pgpu_encrypt_with_r_sk): same ciphertexts, outside the timed steps
        c2_dev = torch.empty_like(c_dev)
        check(lib.pgpu_encrypt_with_r_sk_dev(sk._ctx, count, vp(m_dev), vp(r_dev), vp(c2_dev)), sk._ctx)
        es = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        barrier()
        es[0].record(stream)
        check(lib.pgpu_encrypt_with_r_sk_dev(sk._ctx, count, vp(m_dev), vp(r_dev), vp(c2_dev)), sk._ctx)
        es[1].record(stream)
        barrier()
        assert torch.equal(c2_dev, c_dev), "secret-key EncryptWithR differs from the public-key path"
        breakdown["enc_sk_per_s"] = world * count / (max_over_ranks(es[0].elapsed_time(es[1])) * 1e-3)
        Sk, k_sqr, k_mul = sk.program_cost(3)
        breakdown["enc_sk_program"] = {"limbs": Sk, "sqr": k_sqr, "mul": k_mul, "mac32_per_item": mont_macs(Sk, k_sqr, k_mul)}
        # online half of an offline/online EncryptWithR: c = (1 + m*n) * rn with rn = r^n precomputed (pgpu_encrypt_with_rn)
        check(lib.pgpu_encrypt_with_r_sk_dev(sk._ctx, count, vp(torch.zeros_like(m_dev)), vp(r_dev), vp(c2_dev)), sk._ctx)   # the pool
        c3_dev = torch.empty_like(c_dev)
        check(lib.pgpu_encrypt_with_rn_dev(sk._ctx, count, vp(m_dev), vp(c2_dev), vp(c3_dev)), sk._ctx)
        barrier()
        es[0].record(stream)
        check(lib.pgpu_encrypt_with_rn_dev(sk._ctx, count, vp(m_dev), vp(c2_dev), vp(c3_dev)), sk._ctx)
        es[1].record(stream)
        barrier()
        assert torch.equal(c3_dev, c_dev), "EncryptWithRn(m, r^n) differs from EncryptWithR(m, r)"
        breakdown["enc_online_per_s"] = world * count / (max_over_ranks(es[0].elapsed_time(es[1])) * 1e-3)
        del c2_dev, c3_dev

        # threshold PartialDecrypt at 2048-bit n (BASELINE metric's partial-dec/s), outside the timed steps
        from paillier_b200.keygen import ThresholdKeyGenerator
        tp, tq = synth.load_key("threshold_2048")
        keys = ThresholdKeyGenerator(2048, 8, 5).with_safe_primes(tp, tq).GenerateKeys(device=local)
        tsk = keys[rank % 8]
        for k in keys:
            if k is not tsk:
                k.close()
        check(lib.pgpu_ctx_set_stream(tsk._ctx, C.c_void_p(stream.cuda_stream)), tsk._ctx)
        pcount = max(1, count // 4)
        pin = c_dev[:pcount * w_n2]
        pout = torch.empty_like(pin)
        check(lib.pgpu_partial_decrypt_dev(tsk._ctx, pcount, vp(pin), vp(pout)), tsk._ctx)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        check(lib.pgpu_partial_decrypt_dev(tsk._ctx, pcount, vp(pin), vp(pout)), tsk._ctx)
        e1.record(stream)
        barrier()
        pms = max_over_ranks(e0.elapsed_time(e1))
        Sp, p_sqr, p_mul = tsk.program_cost(2)
        breakdown["pdec_per_s"] = world * pcount / (pms * 1e-3)
        breakdown["pdec_program"] = {"limbs": Sp, "sqr": p_sqr, "mul": p_mul, "mac32_per_item": mont_macs(Sp, p_sqr, p_mul), "items": pcount}
        # PartialDecryptionWithZKP (thresholdkey.go:225-255) on a smaller slice: 3 full exponentiations + a fixed-base one per item
        zcount = max(1, min(pcount, 1 << 14))
        zr = torch.from_numpy(synth.random_records(zcount, w_n2, (n * n).bit_length() - 1, stream=31)).to(dev)
        ze = torch.empty(zcount * 32, dtype=torch.uint8, device=dev)
        zz = torch.empty(zcount * tsk.w_z, dtype=torch.uint8, device=dev)
        zin = c_dev[:zcount * w_n2]
        zout = torch.empty_like(zin)
        for timed in (False, True):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            check(lib.pgpu_pdec_zkp_prove_dev(tsk._ctx, zcount, vp(zin), vp(zr), vp(zout), vp(ze), vp(zz)), tsk._ctx)
            e1.record(stream)
            barrier()
        breakdown["pdec_zkp_prove_per_s"] = world * zcount / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
        breakdown["pdec_zkp_items"] = zcount
        # VerifyProof (thresholdkey.go:278-311) of those proofs
        zok = torch.zeros(zcount, dtype=torch.uint8, device=dev)
        for timed in (False, True):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            check(lib.pgpu_pdec_zkp_verify_dev(tsk._ctx, zcount, tsk.ID, vp(zin), vp(zout), vp(ze), vp(zz), vp(zok)), tsk._ctx)
            e1.record(stream)
            barrier()
        assert bool(zok.all().item()), "ZKP verification rejected an honest proof"
        breakdown["pdec_zkp_verify_per_s"] = world * zcount / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
        tsk.close()
        # BASELINE config 4's kernel: PartialDecrypt at 3072-bit n (6144-bit modulus, powm_vm<8,24>)
        tp3, tq3 = synth.load_key("threshold_3072")
        keys3 = ThresholdKeyGenerator(3072, 8, 5).with_safe_primes(tp3, tq3).GenerateKeys(device=local)
        tsk3 = keys3[rank % 8]
        for k in keys3:
            if k is not tsk3:
                k.close()
        check(lib.pgpu_ctx_set_stream(tsk3._ctx, C.c_void_p(stream.cuda_stream)), tsk3._ctx)
        p3count = max(1, min(count, 1 << 15))
        p3in = torch.from_numpy(synth.random_records(p3count, tsk3.w_n2, 2 * (tp3 * tq3).bit_length() - 2, seed, stream=61)).to(dev)
        p3out = torch.empty_like(p3in)
        for timed in (False, True):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            check(lib.pgpu_partial_decrypt_dev(tsk3._ctx, p3count, vp(p3in), vp(p3out)), tsk3._ctx)
            e1.record(stream)
            barrier()
        S3_, s3q, s3m = tsk3.program_cost(2)
        breakdown["pdec3072_per_s"] = world * p3count / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
        breakdown["pdec3072_program"] = {"limbs": S3_, "sqr": s3q, "mul": s3m, "mac32_per_item": mont_macs(S3_, s3q, s3m), "items": p3count}
        tsk3.close()
        # BASELINE config 3: encrypted dot product with 64-bit scalars (ConstMult + Add) over the ciphertexts of this step
        dcount = count
        k64 = torch.from_numpy(synth.scalars_u64(dcount, seed).view(np.int64).copy()).to(dev)
        dot = torch.empty(w_n2, dtype=torch.uint8, device=dev)
        for timed in (False, True):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            check(lib.pgpu_dot_u64_dev(sk._ctx, dcount, vp(c_dev), C.cast(C.c_void_p(k64.data_ptr()), C.POINTER(C.c_uint64)), vp(dot)), sk._ctx)
            e1.record(stream)
            barrier()
        breakdown["dot_u64_terms_per_s"] = world * dcount / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
        breakdown["dot_u64_terms"] = dcount
        # AltEncryptWithR (paillier.go:221-238): fixed base h_1, comb table, no squarings
        from paillier_b200.api import PublicKey
        hr = int.from_bytes(synth.randomness(1, n, w_n, seed).tobytes(), "little")
        apk = PublicKey(n, device=local, H=hr * hr % n, K=1 << (n.bit_length() // 2))
        check(lib.pgpu_ctx_set_stream(apk._ctx, C.c_void_p(stream.cuda_stream)), apk._ctx)
        acount = max(1, count // 4)
        a_host_m, a_host_r = m_host[:acount * w_n], r_host[:acount * w_n]
        a_host_c = torch.empty(acount * w_n2, dtype=torch.uint8).pin_memory()
        hp2 = lambda t: C.c_void_p(t.data_ptr())
        for timed in (False, True):
            barrier()
            t0 = time.perf_counter()
            check(lib.pgpu_alt_encrypt_with_r_at_level(apk._ctx, 1, acount, hp2(a_host_m), hp2(a_host_r), hp2(a_host_c)), apk._ctx)
            barrier()
            adt = time.perf_counter() - t0
        breakdown["alt_enc_e2e_per_s"] = world * acount / max_over_ranks(adt)
        breakdown["alt_enc_items"] = acount
        apk.close()
        # ---- light operations (1-3 multiplications per record): device-resident rate against the multiplier and HBM roofs,
        # and end to end through the chunked host-buffer ABI (pinned memory) against the measured PCIe rate of this box
        lcount_ = max(1, min(count, 1 << 19))
        pin_big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        dev_big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        pcie = {}
        for name, (dst, src) in (("h2d", (dev_big, pin_big)), ("d2h", (pin_big, dev_big))):
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); dst.copy_(src, non_blocking=True); e1.record(stream); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            pcie[name + "_gbs"] = (256 << 20) / (best * 1e-3) / 1e9
        del pin_big, dev_big
        light = {"items": lcount_, "pcie": pcie}
        rn_dev = torch.empty(lcount_ * w_n2, dtype=torch.uint8, device=dev)
        check(lib.pgpu_encrypt_with_r_sk_dev(sk._ctx, lcount_, vp(torch.zeros(lcount_ * w_n, dtype=torch.uint8, device=dev)), vp(r_dev), vp(rn_dev)), sk._ctx)
        o_dev = torch.empty(lcount_ * w_n2, dtype=torch.uint8, device=dev)
        rn_host = rn_dev.cpu().pin_memory()
        o_host = torch.empty(lcount_ * w_n2, dtype=torch.uint8).pin_memory()
        hpl = lambda t: C.c_void_p(t.data_ptr())
        ops = {
            "add_pairs": (lambda: lib.pgpu_add_pairs_dev(sk._ctx, lcount_, vp(c_dev), vp(rn_dev), vp(o_dev)),
                          lambda: lib.pgpu_add_pairs(sk._ctx, lcount_, hpl(c_host_l), hpl(rn_host), hpl(o_host)), 2 * w_n2, w_n2, 2),
            "encrypt_with_rn": (lambda: lib.pgpu_encrypt_with_rn_dev(sk._ctx, lcount_, vp(m_dev), vp(rn_dev), vp(o_dev)),
                                lambda: lib.pgpu_encrypt_with_rn(sk._ctx, lcount_, hpl(m_host), hpl(rn_host), hpl(o_host)), w_n + w_n2, w_n2, 2),
        }
        c_host_l = c_dev[:lcount_ * w_n2].cpu().pin_memory()
        mulmac = 2.0 * S_ * S_ + S_ if (S_ := sk.program_cost(0)[0]) else 0
        for name, (fdev, fhost, in_b, out_b, nmul) in ops.items():
            for timed in (False, True):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); check(fdev(), sk._ctx); e1.record(stream)
                barrier()
            dms = max_over_ranks(e0.elapsed_time(e1))
            for timed in (False, True):
                barrier()
                t0 = time.perf_counter(); check(fhost(), sk._ctx); hdt = time.perf_counter() - t0
            hdt = max_over_ranks(hdt)
            # the same call on pageable memory (what a Go slice is): copies run at the pageable rate and block the caller
            pg = None
            if name == "add_pairs":
                pa, pb, po = c_host_l.numpy().copy(), rn_host.numpy().copy(), np.empty(lcount_ * w_n2, dtype=np.uint8)
                npv = lambda a: a.ctypes.data_as(C.c_void_p)
                for timed in (False, True):
                    t0 = time.perf_counter(); check(lib.pgpu_add_pairs(sk._ctx, lcount_, npv(pa), npv(pb), npv(po)), sk._ctx); pg = time.perf_counter() - t0
                assert np.array_equal(po, o_host.numpy()), "pageable and pinned host paths differ"
                del pa, pb, po
            bound = max(lcount_ * in_b / (pcie["h2d_gbs"] * 1e9), lcount_ * out_b / (pcie["d2h_gbs"] * 1e9))
            light[name] = {"device_items_per_s": lcount_ / (dms * 1e-3), "device_tmac32": nmul * mulmac * lcount_ / (dms * 1e-3) / 1e12,
                           "device_hbm_gbs": lcount_ * (in_b + out_b) / (dms * 1e-3) / 1e9, "montgomery_muls_per_item": nmul,
                           "e2e_items_per_s": lcount_ / hdt, "e2e_gbs_in": lcount_ * in_b / hdt / 1e9, "e2e_gbs_out": lcount_ * out_b / hdt / 1e9,
                           "e2e_frac_of_pcie_bound": bound / hdt}
            if pg:
                light[name]["e2e_pageable_items_per_s"] = lcount_ / pg
        assert torch.equal(o_host, o_dev.cpu()), "host and device paths of EncryptWithRn differ"
        breakdown["light_ops"] = light
        try:     # AltEncryptWithR end to end (measured above, unchunked host path) against the same PCIe bound
            ab = max(breakdown["alt_enc_items"] * 2 * w_n / (pcie["h2d_gbs"] * 1e9), breakdown["alt_enc_items"] * w_n2 / (pcie["d2h_gbs"] * 1e9))
            light["alt_encrypt"] = {"e2e_items_per_s": breakdown["alt_enc_e2e_per_s"], "e2e_frac_of_pcie_bound": ab / (breakdown["alt_enc_items"] / breakdown["alt_enc_e2e_per_s"] * world),
                                    "note": "~205 fixed-base multiplications per item: bound by the multiplier, not by the link"}
        except Exception:
            pass
        del rn_dev, o_dev, rn_host, o_host, c_host_l
        # level 2 (mod n^3): EncryptWithRAtLevel + Decrypt (CRT over p^3, q^3) through the host-buffer ABI
        lcount = max(1, min(count, 1 << 15))
        l2m = torch.from_numpy(synth.random_records(lcount, w_n2, (n * n).bit_length() - 1, seed, stream=51)).pin_memory()
        l2c = torch.empty(lcount * sk.w_n3, dtype=torch.uint8).pin_memory()
        l2d = torch.empty(lcount * w_n2, dtype=torch.uint8).pin_memory()
        hp3 = lambda t: C.c_void_p(t.data_ptr())
        for timed in (False, True):
            barrier()
            t0 = time.perf_counter()
            check(lib.pgpu_encrypt_with_r_at_level(sk._ctx, 2, lcount, hp3(l2m), hp3(r_host[:lcount * w_n]), hp3(l2c)), sk._ctx)
            barrier()
            t1 = time.perf_counter()
            check(lib.pgpu_decrypt_at_level(sk._ctx, 2, lcount, hp3(l2c), hp3(l2d)), sk._ctx)
            barrier()
            t2 = time.perf_counter()
        assert torch.equal(l2d, l2m), "level 2: Decrypt(Encrypt(m)) != m"
        breakdown["level2_enc_e2e_per_s"] = world * lcount / max_over_ranks(t1 - t0)
        breakdown["level2_dec_e2e_per_s"] = world * lcount / max_over_ranks(t2 - t1)
        breakdown["level2_items"] = lcount
        l2c_sk = torch.empty_like(l2c)                        # the same ciphertexts by the key holder (r^(n^2) over p^3, q^3)
        for timed in (False, True):
            barrier()
            t0 = time.perf_counter()
            check(lib.pgpu_encrypt_with_r_at_level_sk(sk._ctx, 2, lcount, hp3(l2m), hp3(r_host[:lcount * w_n]), hp3(l2c_sk)), sk._ctx)
            barrier()
            t1 = time.perf_counter()
        assert torch.equal(l2c_sk, l2c), "level 2: secret-key EncryptWithRAtLevel differs from the public-key path"
        breakdown["level2_enc_sk_e2e_per_s"] = world * lcount / max_over_ranks(t1 - t0)
        del l2c_sk
        # DDLEQ (ddleq.go): prove + verify, `dsecpar` instances per statement, through the host-buffer ABI on records
        # (marshalling of Python integers stays outside the timed region); device time of the kernels beside it
        if rank == 0:
            from paillier_b200.api import ENC_LEVEL_TWO, to_records
            dn, dsecpar = 2048, 8
            ints = lambda a, w: [int.from_bytes(a[i * w:(i + 1) * w].tobytes(), "little") for i in range(len(a) // w)]
            rr_rec = synth.randomness(dn * (4 + 2 * dsecpar), n, w_n, seed + 7)
            rr = ints(rr_rec, w_n)
            inner = sk.EncryptWithRBatch(ints(synth.plaintexts(dn, n, w_n, seed + 7), w_n), rr[:dn])
            ct1 = sk.EncryptWithRAtLevelBatch([c.C for c in inner], rr[dn:2 * dn], ENC_LEVEL_TWO)
            As, Bs = rr[2 * dn:3 * dn], rr[3 * dn:4 * dn]
            ct2 = sk.NestedRandomizeWithBatch(ct1, As, Bs)
            c1r, c2r = to_records([c.C for c in ct1], sk.w_n3), to_records([c.C for c in ct2], sk.w_n3)
            ar, br = rr_rec[2 * dn * w_n:3 * dn * w_n], rr_rec[3 * dn * w_n:4 * dn * w_n]
            xr = rr_rec[4 * dn * w_n:(4 + dsecpar) * dn * w_n]
            yr = rr_rec[(4 + dsecpar) * dn * w_n:(4 + 2 * dsecpar) * dn * w_n]
            sk.prove_ddleq_records(2, dsecpar, c1r[:2 * sk.w_n3], c2r[:2 * sk.w_n3], ar[:2 * w_n], br[:2 * w_n], xr[:2 * dsecpar * w_n], yr[:2 * dsecpar * w_n])
            check(lib.pgpu_ctx_enable_timing(sk._ctx, 1), sk._ctx)
            kms = C.c_float()
            t0 = time.perf_counter()
            al_r, e_r, f_r = sk.prove_ddleq_records(dn, dsecpar, c1r, c2r, ar, br, xr, yr)
            t1 = time.perf_counter()
            check(lib.pgpu_ctx_last_kernel_ms(sk._ctx, C.byref(kms)), sk._ctx); prove_kms = kms.value
            okd = sk.verify_ddleq_records(dn, dsecpar, c1r, c2r, xr, yr, al_r, e_r, f_r)
            t2 = time.perf_counter()
            check(lib.pgpu_ctx_last_kernel_ms(sk._ctx, C.byref(kms)), sk._ctx); verify_kms = kms.value
            check(lib.pgpu_ctx_enable_timing(sk._ctx, 0), sk._ctx)
            assert bool(okd.all()), "DDLEQ verification rejected an honest proof"
            breakdown["ddleq_prove_instances_per_s"] = dn * dsecpar / (t1 - t0)
            breakdown["ddleq_verify_instances_per_s"] = dn * dsecpar / (t2 - t1)
            if prove_kms > 0 and verify_kms > 0:
                breakdown["ddleq_prove_instances_per_s_device"] = dn * dsecpar / (prove_kms * 1e-3)
                breakdown["ddleq_verify_instances_per_s_device"] = dn * dsecpar / (verify_kms * 1e-3)
            breakdown["ddleq_instances"] = dn * dsecpar
            extras_keep["ddleq"] = (dn, dsecpar, c1r, c2r, xr, yr, al_r, e_r, f_r)
        # BASELINE configs[4]: safe-prime candidate procedure (sieve + Miller-Rabin + Fermat) at 1024-bit p, 2^16 candidates
        # per call, timed around the C-ABI call on byte strings (no conversion to Python integers inside)
        if rank == 0:
            ncand = 1 << 16
            nbc = (1024 - 1 + 7) // 8
            rawb = np.ascontiguousarray(synth.random_records(ncand, nbc, 8 * nbc, stream=41))
            sp_p = np.zeros(ncand * 128, dtype=np.uint8); sp_q = np.zeros(ncand * 128, dtype=np.uint8); sp_ok = np.zeros(ncand, dtype=np.uint8)
            nl = C.c_uint64()
            npp = lambda a: a.ctypes.data_as(C.c_void_p)
            for timed in (False, True):
                cnt = ncand if timed else 1024
                t0 = time.perf_counter()
                rc = lib.pgpu_safe_prime_scan(local, 1024, cnt, npp(rawb), npp(sp_p), npp(sp_q), npp(sp_ok), C.byref(nl))
                spdt = time.perf_counter() - t0
                if rc != 0:
                    raise SystemExit("pgpu_safe_prime_scan failed: " + (lib.pgpu_primes_last_error() or b"").decode())
            breakdown["safe_prime_candidates_per_s"] = ncand / spdt
            breakdown["safe_prime_candidates"] = ncand
            # one 1023-bit Miller-Rabin round = 1023 squarings + ~1023/2 doublings at s = 32 limbs (2.1e6 MAC32);
            # every candidate that survives the sieves (all of them here: the scan tests one survivor per byte string) pays it
            breakdown["safe_prime_mr_mac32_per_candidate"] = mont_macs(32, 1023, 0)
            extras_keep["safe_prime"] = (rawb, sp_ok.copy())

    # ---- strong scaling (configs[1]: "then sharded across 2/4/8"): ONE batch of --strong-count items split over the ranks
    strong = None
    if world > 1 and not args.no_extras:
        from paillier_b200.multi import shard_range
        slo, shi = shard_range(args.strong_count, world, rank)
        scount = min(shi - slo, count)
        def sstep():
            check(lib.pgpu_encrypt_with_r_dev(sk._ctx, scount, vp(m_dev), vp(r_dev), vp(c_dev)), sk._ctx)
            check(lib.pgpu_decrypt_dev(sk._ctx, scount, vp(c_dev), vp(d_dev)), sk._ctx)
        sstep()
        ssteps = 2
        es = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        barrier()
        es[0].record(stream)
        for _ in range(ssteps):
            sstep()
        es[1].record(stream)
        barrier()
        sms_ = max_over_ranks(es[0].elapsed_time(es[1])) / ssteps
        assert torch.equal(d_dev[:scount * w_n], m_dev[:scount * w_n])
        strong = {"value": world * scount / (sms_ * 1e-3), "unit": UNIT, "scaling": "strong", "items_total": world * scount,
                  "items_per_gpu": scount, "ms_per_step": sms_, "steps": ssteps,
                  "note": "one batch split into contiguous slices, no collective on the data path; resident groups per GPU = "
                          "%d, so the last round of the persistent grid is partly idle at small slices" % sk_resident}

    if world == 1 and not args.no_extras and count == args.strong_count:
        strong = {"value": value, "unit": UNIT, "scaling": "strong", "items_total": count, "items_per_gpu": count, "ms_per_step": ms_per_step,
                  "steps": args.steps, "note": "one GPU: the strong-scaling batch is the headline step itself"}

    # ---- Add over a batch sharded across the ranks: per-rank tree product, all-gather of `world` records, one fold
    sharded_add = None
    if world > 1 and not args.no_extras:
        from paillier_b200.multi import sharded_add as _sharded_add
        acount = min(count, 1 << 18)
        part = torch.empty(w_n2, dtype=torch.uint8, device=dev)
        tot = torch.empty(w_n2, dtype=torch.uint8, device=dev)
        def fold(parts):
            check(lib.pgpu_add_reduce_dev(sk._ctx, world, vp(parts), vp(tot)), sk._ctx)
            stream.synchronize()
            return tot
        for timed in (False, True):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            check(lib.pgpu_add_reduce_dev(sk._ctx, acount, vp(c_dev), vp(part)), sk._ctx)
            stream.synchronize()
            total_ct = _sharded_add(dist, rank, world, w_n2, part, fold)
            e1.record(stream)
            barrier()
        ams = max_over_ranks(e0.elapsed_time(e1))
        # every rank decrypts the total; it must be the sum of all ranks' plaintexts mod n
        dsum = torch.empty(w_n, dtype=torch.uint8, device=dev)
        check(lib.pgpu_decrypt_dev(sk._ctx, 1, vp(total_ct), vp(dsum)), sk._ctx)
        stream.synchronize()
        mine_sum = sum(int.from_bytes(m_host.numpy()[i * w_n:(i + 1) * w_n].tobytes(), "little") for i in range(acount)) % n
        sums = torch.zeros(world * w_n, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(sums, torch.frombuffer(bytearray(mine_sum.to_bytes(w_n, "little")), dtype=torch.uint8).to(dev))
        want = sum(int.from_bytes(sums[r_ * w_n:(r_ + 1) * w_n].cpu().numpy().tobytes(), "little") for r_ in range(world)) % n
        got = int.from_bytes(dsum.cpu().numpy().tobytes(), "little")
        assert got == want, "sharded Add: the decrypted total is not the sum of the plaintexts"
        sharded_add = {"ciphertexts_per_gpu": acount, "ms": ams, "ciphertexts_per_s": world * acount / (ams * 1e-3),
                       "exchange": "all-gather of %d records of %d bytes" % (world, w_n2), "decrypts_to_the_sum_of_all_plaintexts": True}

    # ---- BASELINE configs[3]: threshold round, 3072-bit n, 8 shares / threshold 5
    config4 = None
    if not args.no_extras and not args.no_config4 and 8 % world == 0:
        config4 = config4_leg(args, dist, rank, world, local, dev, barrier, max_over_ranks)

    if rank == 0:
        peak = imad_peak() if not args.no_extras else None
        peak_t = peak["imad_wide_tmacs"] if peak else None
        fpeak = fp64_peak() if not args.no_extras else None
        shape_n2, shape_p2 = sk.kernel_shape(1), sk.kernel_shape(3)
        kname = lambda sh, S: ("powm_vm52<%d,%d,%d>" % (sh["tpi"], sh["limbs_per_lane"], S)) if sh["fp64"] else ("powm_vm<%d,%d>" % (sh["tpi"], sh["limbs_per_lane"]))
        achieved = enc_macs * count / (enc_ms * 1e-3) / 1e12      # per GPU: one launch processes `count` items
        # what the multiplier executes: no dedicated squaring in this kernel, every modular multiplication is a full product
        executed = (n_sqr + n_mul) * (2.0 * S * S + S) * count / (enc_ms * 1e-3) / 1e12
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        alg_bytes = count * (2 * w_n + w_n2)
        # DRAM bytes of the EncryptWithR launch: NOT measured in this run -- per-item traffic of the committed `ncu --set full`
        # capture of the same program (window table spilling past the L2) scaled to this launch's item count
        traffic, traffic_src = None, None
        for name in ("r02_ncu_summary.json", "r01_ncu_powm_vm_summary_v4.json"):
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", name)))
                k0 = prof["kernels"][0]
                unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                def _b(sv):
                    v, u = sv.split()[:2]
                    return float(v) * unit[u]
                traffic = (_b(k0["dram__bytes_read.sum"]) + _b(k0["dram__bytes_write.sum"])) / prof["items_per_launch"] * count
                traffic_src = "extrapolated from profiles/%s (%d items per launch there), not measured in this run" % (name, prof["items_per_launch"])
                break
            except Exception:
                continue
        def frac_of(key_prog, key_rate):
            if key_rate not in breakdown:
                return None
            a = breakdown[key_prog]["mac32_per_item"] * breakdown[key_rate] / world / 1e12
            return {"achieved": a, "frac": (a / peak_t) if peak_t else None}
        other = {"crt_decrypt (2 x %s + crt_combine)" % kname(shape_p2, Sd): {
            "achieved": dec_macs * count / (dec_ms * 1e-3) / 1e12,
            "frac": (dec_macs * count / (dec_ms * 1e-3) / 1e12 / peak_t) if peak_t else None}}
        for label, kp, kr in (("partial_decrypt 2048-bit n", "pdec_program", "pdec_per_s"), ("partial_decrypt 3072-bit n", "pdec3072_program", "pdec3072_per_s")):
            f = frac_of(kp, kr)
            if f:
                other[label] = f
        if "safe_prime_candidates_per_s" in breakdown:
            a = breakdown["safe_prime_mr_mac32_per_candidate"] * breakdown["safe_prime_candidates_per_s"] / 1e12
            other["safe_prime_scan (strong_kernel: one Miller-Rabin round per candidate, e2e incl. sieve and copies)"] = {
                "achieved": a, "frac": (a / peak_t) if peak_t else None}
        if "pdec_zkp_prove_per_s" in breakdown and "pdec_program" in breakdown:
            # PartialDecryptionWithZKP = c_i and a = (c^4)^r from one squaring chain (~4112 squarings + 2*(686 + 128) bucket
            # multiplications at 2048-bit n) + b = V^r (fixed base, ~820 multiplications): ~1.5 x the MAC32 of a partial decryption
            pm = breakdown["pdec_program"]["mac32_per_item"]
            a = 1.5 * pm * breakdown["pdec_zkp_prove_per_s"] / world / 1e12
            other["pdec_zkp_prove 2048-bit n (partial decryption + proof, ~1.5 x the partial-decrypt work per item)"] = {
                "achieved": a, "frac": (a / peak_t) if peak_t else None}
        for name in ("add_pairs", "encrypt_with_rn"):
            lo_ = breakdown.get("light_ops", {}).get(name)
            if lo_:
                other["%s (device-resident, %d Montgomery multiplications per record)" % (name, lo_["montgomery_muls_per_item"])] = {
                    "achieved": lo_["device_tmac32"], "frac": (lo_["device_tmac32"] / peak_t) if peak_t else None,
                    "hbm_gbs": lo_["device_hbm_gbs"], "hbm_frac": (lo_["device_hbm_gbs"] / mp["hbm_gbs"]) if mp.get("hbm_gbs") else None}
        roofline = {
            "bound": "imad", "kernel": "%s (EncryptWithR launch)" % kname(shape_n2, S), "achieved": achieved, "peak": peak_t, "unit": "TMAC32/s",
            "frac": (achieved / peak_t) if peak_t else None,
            "peak_source": "tools/imad_peak.cu run live on this GPU: dependency-free IMAD.WIDE.U32 issue rate, the chip's integer-multiply peak "
                           "(SURVEY.md 8d); MEASURED_PEAKS.json holds HBM and bf16 peaks only, which do not bound this carry-chain kernel",
            "mac32_per_item": enc_macs, "items_per_launch": count, "launch_ms": enc_ms,
            "executed_tmac32": executed, "executed_frac": (executed / peak_t) if peak_t else None,
            "note": "achieved counts squarings at 1.5s^2+1.5s and multiplications at 2s^2+s MAC32 (SURVEY.md 8d) whichever pipe executes them; "
                    "executed_* counts every modular multiplication as a full 2s^2+s product, which is what the kernel issues",
            "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (enc_ms * 1e-3) / 1e9,
                    "peak_gbs": mp.get("hbm_gbs"), "frac": (alg_bytes / (enc_ms * 1e-3) / 1e9 / mp["hbm_gbs"]) if mp.get("hbm_gbs") else None},
            "traffic": traffic, "traffic_source": traffic_src,
            "imad_peak": peak,
            "other_kernels": other,
        }
        if shape_n2["fp64"]:
            # the launch runs on the FP64 pipe: 52-bit limbs, 3 FP64 instructions (2 DFMA.RZ + 1 DADD) per limb product,
            # 2*s52^2 limb products per modular multiplication
            s52 = shape_n2["tpi"] * shape_n2["limbs_per_lane"]
            fops = (n_sqr + n_mul) * 2.0 * s52 * s52 * 3.0 * count / (enc_ms * 1e-3) / 1e12
            fp = fpeak["dfma_tops"] if fpeak else None
            roofline["fp64_pipe"] = {"limbs52": s52, "executed_tfp64_inst": fops, "peak_tfp64_inst": fp, "frac": (fops / fp) if fp else None,
                                     "peak_source": "tools/dfma_peak.cu run live: dependency-free DFMA issue rate",
                                     "note": "the kernel alternates FP64 and integer-pipe instructions one to one; ncu shows both pipes "
                                             "throttling each other (profiles/r02_fp64_experiments.md)"}
        cpu = None
        cpu_rows = None
        config1 = None
        if not args.no_extras:
            from oracle import gmp_ref as G
            cores = G.cores()
            sample = args.cpu_sample or 48 * cores
            lam = (p - 1) * (q - 1)
            ms_, rs_ = m_host.numpy()[:sample * w_n], r_host.numpy()[:sample * w_n]
            t0 = time.perf_counter()
            cref = G.encrypt_with_r(n, ms_, rs_, w_n, threads=cores)
            t_enc = time.perf_counter() - t0
            dref = G.decrypt(n, lam, cref, w_n, threads=cores)
            t_all = time.perf_counter() - t0
            # the timed sample doubles as a bit-exact check of the GPU step against the oracle
            assert np.array_equal(cref, c_dev[:sample * w_n2].cpu().numpy()), "GPU ciphertexts differ from the libgmp oracle"
            assert np.array_equal(dref, ms_)
            cpu = {"value": sample / t_all, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {sample} items of rank 0's batch, EncryptWithR + Decrypt as the reference issues them to libgmp "
                             "(paillier.go:213-216, :296-300: two full mpz_powm per encrypt, no CRT); libgmp stand-in for the Go package",
                   "enc_per_s": sample / t_enc, "dec_per_s": sample / (t_all - t_enc)}
            nd = 256 * cores
            c_np = c_dev[:nd * w_n2].cpu().numpy()
            cpu_rows, ref_dot = cpu_baselines_leg(sk, n, p, q, w_n, w_n2, c_np, extras_keep, breakdown)
            config1 = config1_leg(local)
            try:
                breakdown["small_batch"] = small_batch_leg(sk, m_host.numpy(), r_host.numpy(), n, lam, w_n, w_n2)
            except Exception as e:      # a report row, never a reason to lose the bench line
                breakdown["small_batch"] = {"error": repr(e)}
            # the dot-product sample, recomputed on the GPU over the same terms, must equal the libgmp fold
            kd = torch.from_numpy(synth.scalars_u64(nd).view(np.int64).copy()).to(dev)
            dd = torch.empty(w_n2, dtype=torch.uint8, device=dev)
            check(lib.pgpu_dot_u64_dev(sk._ctx, nd, vp(c_dev), C.cast(C.c_void_p(kd.data_ptr()), C.POINTER(C.c_uint64)), vp(dd)), sk._ctx)
            torch.cuda.synchronize(dev)
            assert np.array_equal(ref_dot, dd.cpu().numpy()), "GPU dot product differs from the libgmp oracle"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f64 (52-bit limbs, DFMA) + u64 accumulators" if shape_n2["fp64"] else "u32 limbs (IMAD.WIDE 32x32+64)"), "data": "synthetic",
            "config": {"workload": WORKLOAD, "key": "tests/golden/keys.json:paillier_2048", "items_per_gpu_per_step": count,
                       "items_per_step": world * count,
                       "l2": "inputs per step (%.0f MB) exceed the 126 MB L2" % ((count * (2 * w_n + w_n2)) / 1e6),
                       "sharding": f"{world} independent per-GPU batches, no collective on the data path",
                       "kernels": {"n^2": kname(shape_n2, S), "p^2,q^2": kname(shape_p2, Sd)}},
            "breakdown": breakdown, "roofline": roofline, "cpu_baseline": cpu, "cpu_baselines": cpu_rows, "e2e": e2e, "clocks": clocks,
            "strong": strong, "sharded_add": sharded_add, "config4": config4, "config1": config1,
            "gpu_launches": launches,
        }
        print(json.dumps(line), flush=True)
    sk.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
