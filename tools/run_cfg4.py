#!/usr/bin/env python
"""BASELINE config 4 end to end: a threshold key with `--shares` share-holders spread over the ranks (one per GPU at 8 ranks);
every share-holder computes PartialDecrypt (+ proof with --zkp) for all ciphertexts, NCCL all-gather, each rank verifies
all proofs of its ciphertext slice and combines it.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_cfg4.py --bits 3072 --count 4096 [--zkp]
Checks the recovered plaintexts against the inputs, a slice of the gathered partial decryptions / proofs against the
libgmp oracle on rank 0, and prints one JSON line with device-timed throughput and phases (max over ranks)."""
import argparse
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main_lib(args):
    """the same round through pgpu_multi_threshold_round: one process, one share-holder per visible GPU"""
    import time
    import numpy as np
    import torch
    from oracle import gmp_ref as G
    from paillier_b200 import synth
    from paillier_b200.keygen import ThresholdKeyGenerator
    from paillier_b200.multi import LibThresholdGroup
    n_dev = args.shares
    if torch.cuda.device_count() < n_dev:
        raise SystemExit(f"--impl lib needs {n_dev} visible GPUs (one share-holder per GPU)")
    p, q = synth.load_key(f"threshold_{args.bits}")
    n = p * q
    keys = []
    for d in range(n_dev):
        ks = ThresholdKeyGenerator(args.bits, args.shares, args.threshold, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys(device=d)
        keys.append(ks[d])
        for k in ks:
            if k is not ks[d]:
                k.close()
    t0 = keys[0]
    count = args.count
    m = synth.plaintexts(count, n, t0.w_n)
    c = t0.encrypt_with_r_records(m, synth.randomness(count, n, t0.w_n))
    zr = [synth.random_records(count, t0.w_n2, (n * n).bit_length() - 1, stream=70 + k.ID) for k in keys] if args.zkp else None
    grp = LibThresholdGroup(keys)
    small = min(count, 64)
    grp.round(c[:small * t0.w_n2], [z[:small * t0.w_n2] for z in zr] if zr else None)          # warm-up
    t_0 = time.perf_counter()
    plain, item_ok, phases = grp.round(c, zr)
    dt = time.perf_counter() - t_0
    ok = bool(np.array_equal(plain, m) and item_ok.all())
    par = None
    if args.oracle_items:
        ns = min(count, args.oracle_items)
        ref = G.partial_decrypt(n, t0.Share, args.shares, c[:ns * t0.w_n2], t0.w_n2)
        par = bool(np.array_equal(ref, t0.partial_decrypt_records(c[:ns * t0.w_n2])))
    print(json.dumps({"workload": f"config[3] through pgpu_multi_*: {args.bits}-bit n, {args.shares} shares (threshold {args.threshold}) on {n_dev} GPU(s) "
                                  f"of one process, PartialDecrypt{' + proofs' if args.zkp else ''} over {count} ciphertexts per share-holder",
                      "impl": "lib", "n_gpus": n_dev, "count": count, "ms": dt * 1e3, "ciphertexts_per_s": count / dt,
                      "partial_decryptions_per_s": args.shares * count / dt, "phases_ms": phases,
                      "note": "ms is host wall time of the blocking call, H2D of the ciphertexts and D2H of the plaintexts included",
                      "all_plaintexts_recovered": ok, "oracle_parity": par}), flush=True)
    grp.close()
    for k in keys:
        k.close()
    if not ok or par is False:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=3072, choices=[512, 2048, 3072])
    ap.add_argument("--count", type=int, default=4096)
    ap.add_argument("--shares", type=int, default=8)
    ap.add_argument("--threshold", type=int, default=5)
    ap.add_argument("--zkp", action="store_true")
    ap.add_argument("--oracle-items", type=int, default=16)
    ap.add_argument("--impl", default="torch", choices=["torch", "lib"],
                    help="torch: one process per GPU under torchrun (torch.distributed all-gather); lib: ONE process, pgpu_multi_* "
                         "(ncclCommInitAll inside the library), --shares share-holders on --shares GPUs")
    args = ap.parse_args()
    if args.impl == "lib":
        return main_lib(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    from paillier_b200 import synth
    from paillier_b200.keygen import ThresholdKeyGenerator
    from paillier_b200.multi import gpu_threshold_round_shares

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.shares % world:
        raise SystemExit("--shares must be a multiple of the number of ranks")
    k = args.shares // world
    p, q = synth.load_key(f"threshold_{args.bits}")
    n = p * q
    keys = ThresholdKeyGenerator(args.bits, args.shares, args.threshold, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys(device=local)
    mine = keys[rank * k:(rank + 1) * k]
    for t in keys:
        if t not in mine:
            t.close()
    t0 = mine[0]
    count = args.count
    m = synth.plaintexts(count, n, t0.w_n)
    c = t0.encrypt_with_r_records(m, synth.randomness(count, n, t0.w_n))        # same seeded batch on every rank
    c_dev = torch.from_numpy(c).to(dev)
    zr = [torch.from_numpy(synth.random_records(count, t0.w_n2, (n * n).bit_length() - 1, stream=70 + t.ID)).to(dev) for t in mine] if args.zkp else None
    d = dist if world > 1 else None
    small = min(count, 64)
    gpu_threshold_round_shares(d, mine, c_dev[:small * t0.w_n2], small, world, rank, [z[:small * t0.w_n2] for z in zr] if zr else None)   # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    keep = {}
    plain, (lo, hi), phases = gpu_threshold_round_shares(d, mine, c_dev, count, world, rank, zr, keep=keep)
    e1.record()
    torch.cuda.synchronize()
    ok = bool(np.array_equal(plain.cpu().numpy(), m[lo * t0.w_n:hi * t0.w_n]))
    names = sorted(phases)
    t = torch.tensor([e0.elapsed_time(e1), 0.0 if ok else 1.0] + [phases[x] for x in names], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    par = None
    if rank == 0 and args.oracle_items:
        from oracle import gmp_ref as G
        ns = min(count, args.oracle_items)
        cs = c[:ns * t0.w_n2]
        if args.zkp:
            rd, re_, rz = G.pdec_zkp(n, t0.Share, args.shares, t0.VerificationKey, cs, zr[0][:ns * t0.w_n2].cpu().numpy(), t0.w_n2, t0.w_z)
            par = bool(np.array_equal(rd, keep["dec"][:ns * t0.w_n2].cpu().numpy()) and np.array_equal(re_, keep["e"][:ns * 32].cpu().numpy())
                       and np.array_equal(rz, keep["z"][:ns * t0.w_z].cpu().numpy()))
        else:
            rd = G.partial_decrypt(n, t0.Share, args.shares, cs, t0.w_n2)
            par = bool(np.array_equal(rd, keep["dec"][:ns * t0.w_n2].cpu().numpy()))
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"workload": f"config[3]: {args.bits}-bit n threshold key, {args.shares} shares (threshold {args.threshold}) on {world} GPU(s), "
                                      f"PartialDecrypt{' + proofs' if args.zkp else ''} over {count} ciphertexts per share-holder, all-gather, "
                                      f"{'VerifyProof, ' if args.zkp else ''}sliced Combine",
                          "n_gpus": world, "count": count, "ms": ms, "ciphertexts_per_s": count / (ms * 1e-3),
                          "partial_decryptions_per_s": args.shares * count / (ms * 1e-3),
                          "phases_ms": {x: float(t[2 + i]) for i, x in enumerate(names)},
                          "all_plaintexts_recovered": float(t[1]) == 0.0, "oracle_parity": par}), flush=True)
    for tk in mine:
        tk.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if float(t[1]) != 0.0 or par is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
