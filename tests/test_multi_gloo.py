"""world_size-2 tests of the multi-GPU plumbing (paillier_b200/multi.py) with the gloo backend on CPU
tensors.  The compute steps are injected: here they are the Python oracle on toy keys, which checks the
sharding, the all-gather layout and the share bookkeeping without a GPU."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import paillier_ref as R
from paillier_b200.multi import shard_range, sharded_add, sharded_safe_prime, threshold_round


def test_shard_range_partitions_everything():
    for count in (0, 1, 7, 8, 9, 1000, 2 ** 20 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(count, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


P, Q = 0xe13589576db871aa2575dbb4879698ea1760830894d5e9ef4ee7b7b97a74c287, 0xc2be6b21e328a7f6fdf279ab7df1e39a0d8de7728be82cb3e06b0502894b925f
W2, WN = 128, 64      # record widths for a 512-bit n


def _to(vals, w):
    return torch.frombuffer(bytearray(b"".join(int(v).to_bytes(w, "little") for v in vals)), dtype=torch.uint8).clone()


def _from(t, w):
    b = bytes(t.numpy().tobytes())
    return [int.from_bytes(b[i:i + w], "little") for i in range(0, len(b), w)]


def _worker(rank, world, port, count, drop, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = P * Q
        rnd = random.Random(11)
        nm = n * ((P - 1) // 2) * ((Q - 1) // 2)
        keys = R.threshold_keys_from(P, Q, world, world - (1 if drop is not None else 0), v_seed=rnd.randrange(2, n * n),
                                     coeffs=[rnd.randrange(nm) for _ in range(world - 1 - (1 if drop is not None else 0))])
        pk = R.PublicKey(N=n)
        ms = [rnd.randrange(n) for _ in range(count)]
        cs = [R.encrypt_with_r(pk, m, rnd.randrange(1, n)).C for m in ms]
        tk = R.threshold_public_key(keys[0])

        def partial_decrypt():
            vals = [R.partial_decrypt(keys[rank], c).Decryption for c in cs]
            if drop == rank:                       # a misbehaving server: garbage partials
                vals = [v ^ 1 for v in vals]
            return _to(vals, W2)

        def verify(gathered, r):
            return r != drop

        def combine(gathered, ids, lo, hi):
            rows = gathered.view(world, count, W2)
            out = []
            for i in range(lo, hi):
                shares = [R.PartialDecryption(j, _from(rows[j - 1, i], W2)[0]) for j in ids]
                out.append(R.combine_partial_decryptions(tk, shares))
            return _to(out, WN) if out else torch.zeros(0, dtype=torch.uint8)

        plain, (lo, hi) = threshold_round(dist, rank, world, count, W2, partial_decrypt, combine, verify if drop is not None else None)
        ok = _from(plain, WN) == ms[lo:hi]
        # sharded Add: per-rank product of the slice, gathered and folded
        n2 = n * n
        prod = 1
        for c in cs[lo:hi]:
            prod = prod * c % n2

        def multiply_all(parts):
            acc = 1
            for v in _from(parts, W2):
                acc = acc * v % n2
            return _to([acc], W2)

        total = _from(sharded_add(dist, rank, world, W2, _to([prod], W2), multiply_all), W2)[0]
        ok = ok and total == R.add(pk, *[R.Ciphertext(c) for c in cs]).C
        q.put((rank, ok, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,count,drop", [(2, 5, None), (2, 1, None), (3, 7, 1)])
def test_threshold_round_and_sharded_add_gloo(world, count, drop):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, count, drop, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] == 0 and res[-1][3] == count


# ---- safe-prime search split across ranks (one winner, earliest in stream order) -------------------------

def _oracle_scan(bit_len, raw):
    nb = (bit_len - 1 + 7) // 8
    ps, qs, ok = [], [], []
    for i in range(0, len(raw), nb):
        p, q, good = R.safe_prime_candidate(raw[i:i + nb], bit_len)
        ps.append(p); qs.append(q); ok.append(good)
    return ps, qs, ok


def _stream_reader(seed):
    rnd = random.Random(seed)
    return lambda nbytes: rnd.randbytes(nbytes)


def _sp_worker(rank, world, port, bits, batch, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        reader = _stream_reader(seed) if rank == 0 else None
        q.put((rank, sharded_safe_prime(dist, rank, world, bits, reader, _oracle_scan, batch=batch, max_batches=40)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,bits,batch", [(2, 16, 64), (3, 24, 50)])
def test_sharded_safe_prime_gloo(world, bits, batch):
    seed = 1234
    # single-process answer over the same stream: first accepted candidate in stream order
    want = sharded_safe_prime(None, 0, 1, bits, _stream_reader(seed), _oracle_scan, batch=batch, max_batches=40)
    p, qv = want
    assert p == 2 * qv + 1 and p.bit_length() == bits
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sp_worker, args=(r, world, port, bits, batch, seed, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert [r[1] for r in res] == [want] * world


def test_sharded_safe_prime_gives_up_and_validates_arguments():
    reject_all = lambda bits, raw: ([0] * (len(raw) // 2), [0] * (len(raw) // 2), [False] * (len(raw) // 2))
    with pytest.raises(TimeoutError):                       # safe_prime.go:101-103: the generator gives up
        sharded_safe_prime(None, 0, 1, 16, _stream_reader(1), reject_all, batch=8, max_batches=3)
    with pytest.raises(ValueError):                         # safe_prime.go:67-69
        sharded_safe_prime(None, 0, 1, 5, _stream_reader(1), reject_all)
    with pytest.raises(ValueError):                         # short read from the random source
        sharded_safe_prime(None, 0, 1, 16, lambda nb: b"\x00", reject_all, batch=8, max_batches=1)
