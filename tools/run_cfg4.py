#!/usr/bin/env python
"""BASELINE config 4 end to end: a threshold key with one share-holder per GPU; every rank computes
PartialDecrypt (+ZKP with --zkp) for all ciphertexts, NCCL all-gather, each rank verifies and combines its
slice.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                 --master-port 29511 tools/run_cfg4.py --bits 3072 --count 4096 [--zkp]
Checks the recovered plaintexts against the inputs (and a sample against the oracle on rank 0) and prints one
JSON line with device-timed throughput (max over ranks)."""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=3072, choices=[512, 2048, 3072])
    ap.add_argument("--count", type=int, default=4096)
    ap.add_argument("--zkp", action="store_true")
    ap.add_argument("--threshold", type=int, default=0)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from paillier_b200 import synth
    from paillier_b200.keygen import ThresholdKeyGenerator
    from paillier_b200.multi import gpu_threshold_round

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    p, q = synth.load_key(f"threshold_{args.bits}")
    n = p * q
    w = args.threshold or max(1, (world * 5 + 7) // 8)          # 5 of 8 at world = 8
    keys = ThresholdKeyGenerator(args.bits, world, w, rng=random.Random(5)).with_safe_primes(p, q).GenerateKeys(device=local)
    tsk = keys[rank]
    for k in keys:
        if k is not tsk:
            k.close()
    count = args.count
    m = synth.plaintexts(count, n, tsk.w_n)
    c = tsk.encrypt_with_r_records(m, synth.randomness(count, n, tsk.w_n))        # same seeded batch on every rank
    c_dev = torch.from_numpy(c).to(dev)
    r_dev = torch.from_numpy(synth.random_records(count, tsk.w_n2, (n * n).bit_length() - 1, stream=5)).to(dev) if args.zkp else None
    gpu_threshold_round(dist if world > 1 else None, tsk, c_dev, min(count, 64), world, rank, with_zkp_r=r_dev)   # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plain, (lo, hi) = gpu_threshold_round(dist if world > 1 else None, tsk, c_dev, count, world, rank, with_zkp_r=r_dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = bool(np.array_equal(plain.cpu().numpy(), m[lo * tsk.w_n:hi * tsk.w_n]))
    t = torch.tensor([dt, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"config[3]: {args.bits}-bit n threshold key, {world} shares (threshold {w}), PartialDecrypt"
                                      f"{'+ZKP prove/verify' if args.zkp else ''} over {count} ciphertexts per share-holder, all-gather, sliced Combine",
                          "n_gpus": world, "count": count, "seconds": float(t[0]), "ciphertexts_per_s": count / float(t[0]),
                          "partial_decryptions_per_s": world * count / float(t[0]), "all_plaintexts_recovered": float(t[1]) == 0.0}), flush=True)
    tsk.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if float(t[1]) != 0.0:
        sys.exit(1)


if __name__ == "__main__":
    main()
