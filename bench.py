#!/usr/bin/env python
"""bench.py -- BASELINE.json configs[1]: 2048-bit n, a batch of 2^20 EncryptWithR (fixed, seeded r for the
bit-exact check) followed by CRT Decrypt of the same batch, on N B200s (one process per GPU, the batch
sharded across ranks with no collective on the data path: weak scaling, 2^20 items per GPU).

    python bench.py --gpus N --steps K --warmup W [--count C] [--impl reference]

One "step" = one pass of the hot path over one batch: EncryptWithR over C items, then Decrypt over the C
ciphertexts.  `value` = items through the whole step per second, summed over ranks, inputs resident in HBM.
`e2e` = the same step through the host-buffer C-ABI calls (pgpu_encrypt_with_r / pgpu_decrypt) with pinned
host buffers, H2D and D2H copies inside the timed region.  `breakdown` carries the headline enc/s, dec/s and
partial-dec/s separately.  `--impl reference` times the libgmp restatement of the reference's call sequence
(oracle/gmp_ref.c: a stand-in for the Go package, which cannot be built in this image) on all host cores.
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
        "config": {"workload": WORKLOAD, "key": "tests/golden/keys.json:paillier_2048", "items_per_step": sample},
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


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--count", type=int, default=1 << 20, help="items per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="items of the cpu_baseline sample (default 48 per core)")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu_baseline / partial-decrypt extras")
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
        # DDLEQ (ddleq.go): prove + verify, `dsecpar` instances per statement, through the host-buffer ABI
        if rank == 0:
            from paillier_b200.api import ENC_LEVEL_TWO
            dn, dsecpar = 2048, 8
            ints = lambda a, w: [int.from_bytes(a[i * w:(i + 1) * w].tobytes(), "little") for i in range(len(a) // w)]
            rr = ints(synth.randomness(dn * (4 + 2 * dsecpar), n, w_n, seed + 7), w_n)
            inner = sk.EncryptWithRBatch(ints(synth.plaintexts(dn, n, w_n, seed + 7), w_n), rr[:dn])
            ct1 = sk.EncryptWithRAtLevelBatch([c.C for c in inner], rr[dn:2 * dn], ENC_LEVEL_TWO)
            As, Bs = rr[2 * dn:3 * dn], rr[3 * dn:4 * dn]
            ct2 = sk.NestedRandomizeWithBatch(ct1, As, Bs)
            xs = [rr[4 * dn + i * dsecpar:4 * dn + (i + 1) * dsecpar] for i in range(dn)]
            ys = [rr[(4 + dsecpar) * dn + i * dsecpar:(4 + dsecpar) * dn + (i + 1) * dsecpar] for i in range(dn)]
            sk.ProveDDLEQBatch(dsecpar, ct1[:2], ct2[:2], As[:2], Bs[:2], xs[:2], ys[:2])
            t0 = time.perf_counter()
            proofs = sk.ProveDDLEQBatch(dsecpar, ct1, ct2, As, Bs, xs, ys)
            t1 = time.perf_counter()
            okd = sk.VerifyDDLEQProofBatch(ct1, ct2, proofs)
            t2 = time.perf_counter()
            assert all(okd), "DDLEQ verification rejected an honest proof"
            breakdown["ddleq_prove_instances_per_s"] = dn * dsecpar / (t1 - t0)
            breakdown["ddleq_verify_instances_per_s"] = dn * dsecpar / (t2 - t1)
            breakdown["ddleq_instances"] = dn * dsecpar
        # BASELINE configs[4]: safe-prime candidate procedure (sieve + Miller-Rabin + Fermat) at 1024-bit p
        if rank == 0:
            from paillier_b200.keygen import safe_prime_scan
            ncand = 1 << 15
            rawb = synth.random_records(ncand, 128, 1024, stream=41).tobytes()
            safe_prime_scan(1024, rawb[:128 * 256], device=local)
            t0 = time.perf_counter()
            _, _, okf = safe_prime_scan(1024, rawb, device=local)
            breakdown["safe_prime_candidates_per_s"] = ncand / (time.perf_counter() - t0)
            breakdown["safe_prime_candidates"] = ncand

    if rank == 0:
        peak = imad_peak() if not args.no_extras else None
        peak_t = peak["imad_wide_tmacs"] if peak else None
        achieved = enc_macs * count / (enc_ms * 1e-3) / 1e12      # per GPU: one launch processes `count` items
        # what the multiplier pipe actually executes: powm_vm<4,32> squares with the general multiplier (2s^2+s)
        executed = (n_sqr + n_mul) * (2.0 * S * S + S) * count / (enc_ms * 1e-3) / 1e12
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        alg_bytes = count * (2 * w_n + w_n2)
        # DRAM bytes of the EncryptWithR launch: per-item traffic of the committed `ncu --set full` capture of the same
        # kernel (profiles/r01_ncu_powm_vm_summary_v4.json, 18944 items) scaled to this launch's item count
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_powm_vm_summary_v4.json")))
            k0 = prof["kernels"][0]
            unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            def _b(sv):
                v, u = sv.split()[:2]
                return float(v) * unit[u]
            traffic = (_b(k0["dram__bytes_read.sum"]) + _b(k0["dram__bytes_write.sum"])) / prof["items_per_launch"] * count
        except Exception:
            traffic = None
        roofline = {
            "bound": "imad", "kernel": "powm_vm (EncryptWithR launch)", "achieved": achieved, "peak": peak_t, "unit": "TMAC32/s",
            "frac": (achieved / peak_t) if peak_t else None,
            "peak_source": "tools/imad_peak.cu run live on this GPU: dependency-free IMAD.WIDE.U32 issue rate (SURVEY.md 8d); "
                           "MEASURED_PEAKS.json holds HBM and bf16 peaks only, which do not bound this integer carry-chain kernel",
            "mac32_per_item": enc_macs, "items_per_launch": count, "launch_ms": enc_ms,
            "executed_tmac32": executed, "executed_frac": (executed / peak_t) if peak_t else None,
            "note": "achieved counts squarings at 1.5s^2+1.5s (SURVEY.md 8d); the 4096-bit kernel issues 2s^2+s for them, see executed_*",
            "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (enc_ms * 1e-3) / 1e9,
                    "peak_gbs": mp.get("hbm_gbs"), "frac": (alg_bytes / (enc_ms * 1e-3) / 1e9 / mp["hbm_gbs"]) if mp.get("hbm_gbs") else None},
            "traffic": traffic,
            "imad_peak": peak,
            "other_kernels": {
                "crt_decrypt (2 x powm_vm<4,16> + crt_combine)": {
                    "achieved": dec_macs * count / (dec_ms * 1e-3) / 1e12,
                    "frac": (dec_macs * count / (dec_ms * 1e-3) / 1e12 / peak_t) if peak_t else None},
                **({"partial_decrypt (powm_vm<4,32>)": {
                    "achieved": breakdown["pdec_program"]["mac32_per_item"] * breakdown["pdec_per_s"] / world / 1e12,
                    "frac": (breakdown["pdec_program"]["mac32_per_item"] * breakdown["pdec_per_s"] / world / 1e12 / peak_t) if peak_t else None}}
                   if "pdec_per_s" in breakdown else {}),
                **({"partial_decrypt 3072-bit n (powm_vm<8,24>)": {
                    "achieved": breakdown["pdec3072_program"]["mac32_per_item"] * breakdown["pdec3072_per_s"] / world / 1e12,
                    "frac": (breakdown["pdec3072_program"]["mac32_per_item"] * breakdown["pdec3072_per_s"] / world / 1e12 / peak_t) if peak_t else None}}
                   if "pdec3072_per_s" in breakdown else {}),
            },
        }
        cpu = None
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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (IMAD.WIDE 32x32+64)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "key": "tests/golden/keys.json:paillier_2048", "items_per_gpu_per_step": count,
                       "l2": "inputs per step (%.0f MB) exceed the 126 MB L2" % ((count * (2 * w_n + w_n2)) / 1e6),
                       "sharding": f"{world} independent per-GPU batches, no collective on the data path"},
            "breakdown": breakdown, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches,
        }
        print(json.dumps(line), flush=True)
    sk.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
