"""Multi-GPU plumbing: one process per GPU (torch.distributed), SURVEY.md 8(e).

* Encrypt / Decrypt / ConstMult / proofs / safe-prime candidates are independent items: `shard_range`
  gives rank r its contiguous slice of a batch; there is NO collective on that path.
* Add over a sharded batch: every rank reduces its slice to one ciphertext, the `world` partial products
  are gathered and multiplied (modular multiplication is associative and commutative and the result is a
  canonical residue, so the tree order is bit-identical to the reference's left fold, operations.go:11-29).
* Threshold decryption with one share-holder per GPU (BASELINE config 4): rank r holds share r+1, computes
  the partial decryptions of ALL ciphertexts, one all-gather moves them ([share][ciphertext] layout), then
  rank r combines ciphertext slice r (thresholdkey.go:149-161) straight out of the gathered buffer.

* Safe-prime search (threshold key generation, SURVEY 8e "replicas only, one winner"): rank 0 reads a batch of
  candidate byte strings from the random source and broadcasts it, rank r tests slice r, one MIN all-reduce picks the
  earliest accepted candidate in stream order -- the same (p, q) a single device returns for that stream.

The compute steps are injected as callables so that the plumbing can be exercised with the gloo backend on
CPU tensors in tests (tests/test_multi_gloo.py); `gpu_threshold_round_shares` is the same round on the CUDA engine.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional, Sequence, Tuple


def shard_range(count: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of `count` items for `rank`; the first count % world ranks get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(count, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def threshold_round(dist, rank: int, world: int, count: int, record_width: int,
                    partial_decrypt: Callable[[], "torch.Tensor"],
                    combine: Callable[["torch.Tensor", Sequence[int], int, int], "torch.Tensor"],
                    verify: Optional[Callable[["torch.Tensor", int], bool]] = None,
                    verify_all: Optional[Callable[["torch.Tensor"], Sequence[bool]]] = None):
    """One threshold-decryption round.

    partial_decrypt() -> uint8 tensor [count * record_width]: this rank's share applied to every ciphertext.
    combine(gathered, ids, lo, hi) -> plaintext records of ciphertexts [lo, hi) from the gathered
        [world][count][record_width] buffer using the shares `ids` (server ids, rank r holds id r+1).
    verify(gathered, r) (optional) -> whether rank r's partials are to be used at all.
    verify_all(gathered) (optional) -> one verdict per rank, all shares checked in one batch (takes precedence).
    These two hooks drop a share as a whole: they exist for the plumbing tests on CPU tensors.  The CUDA drivers below do not use
    them -- they filter per ciphertext, as CombinePartialDecryptionsZKP does (thresholdkey.go:164-172), inside
    `gpu_threshold_round_shares` (pgpu_combine_verified_dev).
    Returns (plaintext slice tensor, (lo, hi))."""
    import torch
    mine = partial_decrypt()
    if mine.numel() != count * record_width:
        raise ValueError("partial_decrypt returned a buffer of the wrong size")
    gathered = torch.empty(world * count * record_width, dtype=torch.uint8, device=mine.device)
    if world > 1:
        dist.all_gather_into_tensor(gathered, mine.contiguous())
    else:
        gathered.copy_(mine)
    if verify_all is not None:
        verdicts = list(verify_all(gathered))
        ids = [r + 1 for r in range(world) if verdicts[r]]
    else:
        ids = [r + 1 for r in range(world) if verify is None or verify(gathered, r)]
    lo, hi = shard_range(count, world, rank)
    return combine(gathered, ids, lo, hi), (lo, hi)


def sharded_add(dist, rank: int, world: int, record_width: int, local_product: "torch.Tensor",
                multiply_all: Callable[["torch.Tensor"], "torch.Tensor"]):
    """Add over a batch sharded across ranks: gather the `world` per-rank products (one record each) and fold them."""
    import torch
    parts = torch.empty(world * record_width, dtype=torch.uint8, device=local_product.device)
    if world > 1:
        dist.all_gather_into_tensor(parts, local_product.contiguous())
    else:
        parts.copy_(local_product)
    return multiply_all(parts)


def sharded_safe_prime(dist, rank: int, world: int, bit_len: int, random_reader: Optional[Callable[[int], bytes]],
                       scan: Callable[[int, bytes], Tuple[Sequence[int], Sequence[int], Sequence[bool]]],
                       batch: int = 1 << 14, max_batches: int = 64):
    """GenerateSafePrime (safe_prime.go:61-105) with the candidate stream split across ranks.

    random_reader(nbytes) -> bytes is consulted on rank 0 only (other ranks may pass None); every batch is broadcast so
    that all ranks see the same stream.  scan(bit_len, raw) -> (ps, qs, accepted) is the per-candidate procedure
    (keygen.safe_prime_scan on a GPU; a CPU stand-in in the gloo tests).  Returns (p, q) of the first accepted
    candidate in stream order on every rank -- identical to the single-device search over the same stream."""
    import torch
    if bit_len < 6:
        raise ValueError("safe prime size must be at least 6 bits")               # safe_prime.go:67-69
    nb = (bit_len - 1 + 7) // 8
    pw = (bit_len + 7) // 8
    NONE = 1 << 62
    for _ in range(max_batches):
        raw = torch.empty(batch * nb, dtype=torch.uint8)
        if rank == 0:
            data = random_reader(batch * nb)
            if len(data) != batch * nb:
                raise ValueError("random source returned a short read")
            raw.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        if world > 1:
            dist.broadcast(raw, src=0)
        lo, hi = shard_range(batch, world, rank)
        ps, qs, ok = scan(bit_len, raw[lo * nb:hi * nb].numpy().tobytes()) if hi > lo else ([], [], [])
        first = next((i for i, good in enumerate(ok) if good), None)
        best = torch.tensor([NONE if first is None else lo + first], dtype=torch.int64)
        if world > 1:
            dist.all_reduce(best, op=dist.ReduceOp.MIN)
        win = int(best.item())
        if win == NONE:
            continue
        owner = next(r for r in range(world) if shard_range(batch, world, r)[0] <= win < shard_range(batch, world, r)[1])
        pq = torch.zeros(2 * pw, dtype=torch.uint8)
        if rank == owner:
            pq.copy_(torch.frombuffer(bytearray(ps[win - lo].to_bytes(pw, "big") + qs[win - lo].to_bytes(pw, "big")), dtype=torch.uint8))
        if world > 1:
            dist.broadcast(pq, src=owner)
        b = pq.numpy().tobytes()
        return int.from_bytes(b[:pw], "big"), int.from_bytes(b[pw:], "big")
    raise TimeoutError(f"generator gave up after {max_batches} batches of {batch} candidates")   # safe_prime.go:101-103


def gpu_safe_prime_search(dist, rank: int, world: int, bit_len: int, random_reader, device: int, batch: int = 1 << 14, max_batches: int = 64):
    """sharded_safe_prime bound to the CUDA candidate procedure (keygen.safe_prime_scan) on `device`"""
    from .keygen import safe_prime_scan
    return sharded_safe_prime(dist, rank, world, bit_len, random_reader, lambda bits, raw: safe_prime_scan(bits, raw, device),
                              batch=batch, max_batches=max_batches)


def gpu_threshold_round(dist, tsk, c_dev, count: int, world: int, rank: int, with_zkp_r=None):
    """BASELINE config 4 with ONE share-holder per rank: tsk = this rank's ThresholdSecretKey (share id rank+1), c_dev = device
    tensor of `count` n2-width ciphertext records (the same on every rank), with_zkp_r = optional device tensor of count
    n2-width r values (proofs are produced, all-gathered and verified before the slice is combined).  It is
    `gpu_threshold_round_shares` with a single local share, so the proofs filter PER CIPHERTEXT like the reference's
    CombinePartialDecryptionsZKP (thresholdkey.go:164-172).  Returns (plaintext slice tensor, (lo, hi))."""
    out, span, _ = gpu_threshold_round_shares(dist, [tsk], c_dev, count, world, rank, None if with_zkp_r is None else [with_zkp_r])
    return out, span


def gpu_threshold_round_shares(dist, tsks, c_dev, count: int, world: int, rank: int, zkp_r=None, keep=None):
    """BASELINE config 4 with a fixed number of share-holders spread over `world` GPUs: `tsks` = this rank's
    ThresholdSecretKeys (share ids rank*k+1 .. rank*k+k, k = len(tsks), the same k on every rank; 8 shares on 8 GPUs = one
    share-holder per GPU).  Every share-holder computes PartialDecrypt (thresholdkey.go:192-201) and, with `zkp_r` (one
    device tensor of count n2-width r records per local share), the proof (thresholdkey.go:225-255) for ALL `count`
    ciphertexts; one all-gather per field lands [share][ciphertext]; this rank then verifies every share's proofs of
    its ciphertext slice (VerifyProof, thresholdkey.go:278-311) and combines the slice from the shares that verified
    (CombinePartialDecryptionsZKP, thresholdkey.go:164-172).

    Returns (plaintext slice tensor, (lo, hi), phases) with phases = device milliseconds of this rank per phase
    {"pdec", "prove", "all_gather", "verify", "combine"} (CUDA events on the stream the work is enqueued on).
    keep: optional dict that receives the gathered buffers ("dec", "e", "z") for parity checks."""
    import torch
    from ._lib import check, lib
    k = len(tsks)
    t0 = tsks[0]
    w2, wn, wz = t0.w_n2, t0.w_n, t0.w_z
    dev = c_dev.device
    vp = lambda t: C.c_void_p(t.data_ptr())
    stream = torch.cuda.Stream(dev)
    stream.wait_stream(torch.cuda.current_stream(dev))
    for t in tsks:
        check(lib.pgpu_ctx_set_stream(t._ctx, C.c_void_p(stream.cuda_stream)), t._ctx)
    ev = {name: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for name in ("pdec", "prove", "all_gather", "verify", "combine")}
    shares = world * k
    lo, hi = shard_range(count, world, rank)
    n = hi - lo
    try:
        with torch.cuda.stream(stream):
            dec = torch.empty(k * count * w2, dtype=torch.uint8, device=dev)
            e = z = None
            fused = zkp_r is not None and not os.environ.get("PGPU_NO_FUSED_PROVE")
            ev["pdec"][0].record(stream)
            if not fused:
                for j, t in enumerate(tsks):
                    check(lib.pgpu_partial_decrypt_dev(t._ctx, count, vp(c_dev), vp(dec[j * count * w2:])), t._ctx)
            ev["pdec"][1].record(stream)
            ev["prove"][0].record(stream)
            if zkp_r is not None:
                e = torch.empty(k * count * 32, dtype=torch.uint8, device=dev)
                z = torch.empty(k * count * wz, dtype=torch.uint8, device=dev)
                for j, t in enumerate(tsks):
                    # (slices of live tensors: their storage stays allocated)
                    if fused:      # PartialDecrypt and (c^4)^r in one launch with shared squarings: the "pdec" phase is empty
                        check(lib.pgpu_pdec_zkp_prove_dev(t._ctx, count, vp(c_dev), vp(zkp_r[j]), vp(dec[j * count * w2:]),
                                                          vp(e[j * count * 32:]), vp(z[j * count * wz:])), t._ctx)
                    else:
                        check(lib.pgpu_pdec_zkp_prove_given_dev(t._ctx, count, vp(c_dev), vp(zkp_r[j]), vp(dec[j * count * w2:]),
                                                                vp(e[j * count * 32:]), vp(z[j * count * wz:])), t._ctx)
            ev["prove"][1].record(stream)
            # ---- the one exchange of the path: [share][ciphertext] on every rank
            ev["all_gather"][0].record(stream)
            if world > 1:
                g_dec = torch.empty(shares * count * w2, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(g_dec, dec)
                if zkp_r is not None:
                    g_e = torch.empty(shares * count * 32, dtype=torch.uint8, device=dev)
                    g_z = torch.empty(shares * count * wz, dtype=torch.uint8, device=dev)
                    dist.all_gather_into_tensor(g_e, e)
                    dist.all_gather_into_tensor(g_z, z)
            else:
                g_dec, g_e, g_z = dec, e, z
            ev["all_gather"][1].record(stream)
            # ---- VerifyProof of every share for this rank's ciphertext slice
            ids = list(range(1, shares + 1))
            ok = None
            ev["verify"][0].record(stream)
            if zkp_r is not None and n > 0:
                idarr = (C.c_int * shares)(*ids)
                if shares <= 8 and not os.environ.get("PGPU_NO_SHARED_VERIFY"):
                    # the k proofs of one ciphertext raise the same c^4 to k different Z: item-major records, shared squarings
                    im = lambda buf, w: buf.view(shares, count, w)[:, lo:hi].transpose(0, 1).contiguous().view(-1)
                    v_c, v_dec, v_e, v_z = c_dev[lo * w2:hi * w2], im(g_dec, w2), im(g_e, 32), im(g_z, wz)
                    ok_im = torch.zeros(n * shares, dtype=torch.uint8, device=dev)
                    check(lib.pgpu_pdec_zkp_verify_shared_dev(t0._ctx, n, shares, idarr, vp(v_c), vp(v_dec), vp(v_e), vp(v_z), vp(ok_im)), t0._ctx)
                    ok = ok_im.view(n, shares).t().contiguous().view(-1)           # share-major for the combiner
                else:
                    rows = lambda buf, w: torch.cat([buf[(s * count + lo) * w:(s * count + hi) * w] for s in range(shares)])
                    ok = torch.zeros(shares * n, dtype=torch.uint8, device=dev)
                    # (named tensors: a temporary would hand its memory back to torch's allocator before the kernels ran)
                    v_c, v_dec, v_e, v_z = c_dev[lo * w2:hi * w2].repeat(shares), rows(g_dec, w2), rows(g_e, 32), rows(g_z, wz)
                    check(lib.pgpu_pdec_zkp_verify_multi_dev(t0._ctx, n, shares, idarr, vp(v_c), vp(v_dec), vp(v_e), vp(v_z), vp(ok)), t0._ctx)
            ev["verify"][1].record(stream)
            # ---- Combine the slice in place out of the gathered buffer; with proofs the reference's per-ciphertext filter
            # (CombinePartialDecryptionsZKP, thresholdkey.go:164-172) picks the shares whose proof holds
            out = torch.empty(max(n, 1) * wn, dtype=torch.uint8, device=dev)
            item_ok = torch.ones(max(n, 1), dtype=torch.uint8, device=dev)
            idarr = (C.c_int * shares)(*ids)
            ev["combine"][0].record(stream)
            if n > 0:
                base = g_dec[lo * w2:]
                if ok is None:
                    check(lib.pgpu_combine_strided_dev(t0._ctx, n, shares, idarr, vp(base), count, vp(out)), t0._ctx)
                else:
                    rc = lib.pgpu_combine_verified_dev(t0._ctx, n, shares, idarr, vp(base), count, vp(ok), vp(out), vp(item_ok))
                    if rc not in (0, 6):              # PGPU_ERR_THRESHOLD: some ciphertexts lack valid shares, flagged in item_ok
                        check(rc, t0._ctx)
            ev["combine"][1].record(stream)
        stream.synchronize()
        if keep is not None:
            keep["dec"], keep["e"], keep["z"], keep["ids"] = g_dec, g_e if zkp_r is not None else None, g_z if zkp_r is not None else None, ids
            keep["ok"], keep["item_ok"] = ok, item_ok[:n]
    finally:
        for t in tsks:
            check(lib.pgpu_ctx_set_stream(t._ctx, None), t._ctx)
    phases = {name: a.elapsed_time(b) for name, (a, b) in ev.items()}
    return out[:n * wn], (lo, hi), phases


class LibThresholdGroup:
    """pgpu_multi_* (csrc/multi.cu): the threshold round of BASELINE config 4 inside the library, one ThresholdSecretKey
    context per device of THIS process, NCCL communicators made with ncclCommInitAll.  This is what the Go drop-in calls
    (one cgo call per round); `gpu_threshold_round_shares` above is the one-process-per-GPU form used under torchrun."""

    def __init__(self, tsks):
        from ._lib import check, lib
        self.tsks = list(tsks)
        arr = (C.c_void_p * len(self.tsks))(*[t._ctx for t in self.tsks])
        self._h = C.c_void_p()
        check(lib.pgpu_multi_create(C.byref(self._h), arr, len(self.tsks)))

    def round(self, c, zkp_r=None):
        """c: count n2-width ciphertext records (numpy uint8); zkp_r: one array of count n2-width r records per share-holder
        or None -> (plaintext records, per-ciphertext flags, device milliseconds per phase)"""
        import numpy as np
        from ._lib import PGPU_ERR_THRESHOLD, PGPU_OK, PgpuError, lib
        t0 = self.tsks[0]
        c = np.ascontiguousarray(c).view(np.uint8).reshape(-1)
        count = c.size // t0.w_n2
        plain = np.zeros(count * t0.w_n, dtype=np.uint8)
        item_ok = np.zeros(count, dtype=np.uint8)
        rp = None
        if zkp_r is not None:
            zkp_r = [np.ascontiguousarray(r).view(np.uint8).reshape(-1) for r in zkp_r]
            if len(zkp_r) != len(self.tsks) or any(r.size != count * t0.w_n2 for r in zkp_r):
                raise ValueError("one array of count n2-width records per share-holder")
            rp = (C.c_void_p * len(zkp_r))(*[r.ctypes.data for r in zkp_r])
        rc = lib.pgpu_multi_threshold_round(self._h, count, c.ctypes.data_as(C.c_void_p), rp, plain.ctypes.data_as(C.c_void_p),
                                            item_ok.ctypes.data_as(C.c_void_p))
        if rc not in (PGPU_OK, PGPU_ERR_THRESHOLD):
            raise PgpuError(rc, (lib.pgpu_multi_last_error(self._h) or b"").decode("utf-8", "replace"))
        ph = (C.c_float * 5)()
        lib.pgpu_multi_last_phases_ms(self._h, ph)
        return plain, item_ok, dict(zip(("pdec", "prove", "all_gather", "verify", "combine"), [float(x) for x in ph]))

    def close(self):
        from ._lib import lib
        if self._h:
            lib.pgpu_multi_destroy(self._h)
            self._h = C.c_void_p()
