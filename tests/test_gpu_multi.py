"""GPU side of the multi-GPU plumbing: gpu_threshold_round on one device (world = 1: a 1-of-1 threshold key),
and, when launched under torchrun with >= 2 GPUs (tools/run_cfg4.py), the all-gather path."""
import random

import pytest
import torch

from paillier_b200 import synth
from paillier_b200.api import from_records
from paillier_b200.keygen import ThresholdKeyGenerator
from paillier_b200.multi import gpu_threshold_round

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("zkp", [False, True])
def test_threshold_round_world1(zkp):
    p, q = synth.load_key("threshold_512")
    n = p * q
    tsk = ThresholdKeyGenerator(512, 1, 1, rng=random.Random(4)).with_safe_primes(p, q).GenerateKeys()[0]
    count = 33
    m = synth.plaintexts(count, n, tsk.w_n)
    c = tsk.encrypt_with_r_records(m, synth.randomness(count, n, tsk.w_n))
    dev = torch.device("cuda", 0)
    c_dev = torch.from_numpy(c).to(dev)
    r_dev = torch.from_numpy(synth.random_records(count, tsk.w_n2, (n * n).bit_length() - 1, stream=5)).to(dev) if zkp else None
    plain, (lo, hi) = gpu_threshold_round(None, tsk, c_dev, count, 1, 0, with_zkp_r=r_dev)
    assert (lo, hi) == (0, count)
    assert from_records(plain.cpu().numpy(), tsk.w_n) == from_records(m, tsk.w_n)
    tsk.close()


def test_device_resident_chain():
    # Encrypt -> Randomize -> ConstMult -> Add/Sub pairs -> Add over the batch -> Decrypt without leaving the GPU
    import ctypes as C
    import numpy as np
    from paillier_b200._lib import check, lib
    from paillier_b200.api import SecretKey
    p, q = synth.load_key("paillier_2048")
    n = p * q
    sk = SecretKey(n, p=p, q=q)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    check(lib.pgpu_ctx_set_stream(sk._ctx, C.c_void_p(stream.cuda_stream)), sk._ctx)
    count = 2000
    m = synth.plaintexts(count, n, sk.w_n)
    k = synth.scalars_u64(count)
    vp = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.stream(stream):
        md = torch.from_numpy(m).to(dev)
        rd = torch.from_numpy(synth.randomness(count, n, sk.w_n)).to(dev)
        r2 = torch.from_numpy(synth.randomness(count, n, sk.w_n, synth.SEED + 1)).to(dev)
        kd = torch.from_numpy(k.view(np.int64).copy()).to(dev)
        c = torch.empty(count * sk.w_n2, dtype=torch.uint8, device=dev)
        c2, c3, c4 = torch.empty_like(c), torch.empty_like(c), torch.empty_like(c)
        tot = torch.empty(sk.w_n2, dtype=torch.uint8, device=dev)
        out = torch.empty(sk.w_n, dtype=torch.uint8, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.pgpu_encrypt_with_r_dev(sk._ctx, count, vp(md), vp(rd), vp(c)), sk._ctx)
        check(lib.pgpu_randomize_with_r_dev(sk._ctx, count, vp(c), vp(r2), vp(c2)), sk._ctx)          # same plaintexts, fresh randomness
        check(lib.pgpu_const_mult_dev(sk._ctx, count, vp(c2), vp(kd), 8, vp(c3)), sk._ctx)            # k_i * m_i
        check(lib.pgpu_add_pairs_dev(sk._ctx, count, vp(c3), vp(c), vp(c4)), sk._ctx)                 # + m_i
        check(lib.pgpu_sub_pairs_dev(sk._ctx, count, vp(c4), vp(c2), vp(c4), vp(bad)), sk._ctx)       # - m_i
        check(lib.pgpu_add_reduce_dev(sk._ctx, count, vp(c4), vp(tot)), sk._ctx)
        check(lib.pgpu_decrypt_dev(sk._ctx, 1, vp(tot), vp(out)), sk._ctx)
    stream.synchronize()
    assert int(bad.item()) == -1
    ms = from_records(m, sk.w_n)
    expect = sum(int(ki) * mi for ki, mi in zip(k, ms)) % n
    assert from_records(out.cpu().numpy(), sk.w_n) == [expect]
    check(lib.pgpu_ctx_set_stream(sk._ctx, None), sk._ctx)
    sk.close()


def test_sharded_safe_prime_world1_matches_single_device_search():
    # SURVEY 8(e), threshold keygen: the split search returns the first accepted candidate in stream order,
    # i.e. what GenerateSafePrime returns for the same stream
    from paillier_b200.keygen import GenerateSafePrime, safe_prime_scan
    from paillier_b200.multi import sharded_safe_prime
    for bits, batch in ((64, 512), (128, 2048)):
        r1, r2 = random.Random(77), random.Random(77)
        want = GenerateSafePrime(bits, lambda nb: r1.randbytes(nb), batch=batch)
        got = sharded_safe_prime(None, 0, 1, bits, lambda nb: r2.randbytes(nb), safe_prime_scan, batch=batch)
        assert got == want and got[0] == 2 * got[1] + 1 and got[0].bit_length() == bits


@pytest.mark.parametrize("bits,l,w", [(512, 4, 3), (2048, 8, 5)])
def test_threshold_round_all_shares_on_one_device(bits, l, w):
    # BASELINE config 4 with every share-holder on this device (bench.py's config4 leg at N = 1): PartialDecrypt, proof
    # with the partial decryptions given, VerifyProof of all shares, Combine; per-share output equals the one-call path
    import numpy as np
    from paillier_b200.multi import gpu_threshold_round_shares
    p, q = synth.load_key(f"threshold_{bits}")
    n = p * q
    keys = ThresholdKeyGenerator(bits, l, w, rng=random.Random(11)).with_safe_primes(p, q).GenerateKeys()
    t0 = keys[0]
    count = 21
    m = synth.plaintexts(count, n, t0.w_n)
    c = t0.encrypt_with_r_records(m, synth.randomness(count, n, t0.w_n))
    dev = torch.device("cuda", 0)
    c_dev = torch.from_numpy(c).to(dev)
    rs = [synth.random_records(count, t0.w_n2, (n * n).bit_length() - 1, stream=5 + t.ID) for t in keys]
    keep = {}
    plain, (lo, hi), phases = gpu_threshold_round_shares(None, keys, c_dev, count, 1, 0, [torch.from_numpy(r).to(dev) for r in rs], keep=keep)
    assert (lo, hi) == (0, count) and keep["ids"] == list(range(1, l + 1))
    assert from_records(plain.cpu().numpy(), t0.w_n) == from_records(m, t0.w_n)
    assert set(phases) == {"pdec", "prove", "all_gather", "verify", "combine"} and all(v >= 0 for v in phases.values())
    for j, t in enumerate(keys):
        dec, e, z = t.zkp_prove_records(c, rs[j])
        assert np.array_equal(dec, keep["dec"][j * count * t.w_n2:(j + 1) * count * t.w_n2].cpu().numpy())
        assert np.array_equal(e, keep["e"][j * count * 32:(j + 1) * count * 32].cpu().numpy())
        assert np.array_equal(z, keep["z"][j * count * t.w_z:(j + 1) * count * t.w_z].cpu().numpy())
    # without proofs
    plain2, _, _ = gpu_threshold_round_shares(None, keys, c_dev, count, 1, 0, None)
    assert torch.equal(plain2, plain)
    for t in keys:
        t.close()


@pytest.mark.parametrize("zkp", [False, True])
def test_threshold_round_world2_under_torchrun(zkp):
    # the all-gather path of config 4 on two GPUs (NCCL), 8 share-holders, 4 per rank; skipped on a single-GPU box
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "run_cfg4.py"), "--bits", "512", "--count", "301", "--oracle-items", "301"]
    r = subprocess.run(cmd + (["--zkp"] if zkp else []), capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["all_plaintexts_recovered"] and line["oracle_parity"] is True


def _lib_group(bits, n_dev, l, w):
    import random as _r
    from paillier_b200.multi import LibThresholdGroup
    p, q = synth.load_key(f"threshold_{bits}")
    keys = []
    for d in range(n_dev):
        ks = ThresholdKeyGenerator(bits, l, w, rng=_r.Random(13)).with_safe_primes(p, q).GenerateKeys(device=d)
        keys.append(ks[d])
        for k in ks:
            if k is not ks[d]:
                k.close()
    return p * q, keys, LibThresholdGroup(keys)


@pytest.mark.parametrize("zkp", [False, True])
def test_library_threshold_round(zkp):
    # pgpu_multi_* (csrc/multi.cu): single-process ncclCommInitAll over the visible devices, one share-holder per device.
    # On a one-GPU box this is a 1-of-1 key with a one-rank communicator; with more GPUs the all-gather crosses NVLink.
    import numpy as np
    from oracle import gmp_ref as G
    n_dev = min(torch.cuda.device_count(), 8)
    w = max(1, (5 * n_dev + 7) // 8)
    n, keys, grp = _lib_group(512, n_dev, n_dev, w)
    t0 = keys[0]
    count = 257
    m = synth.plaintexts(count, n, t0.w_n)
    c = t0.encrypt_with_r_records(m, synth.randomness(count, n, t0.w_n))
    rs = [synth.random_records(count, t0.w_n2, (n * n).bit_length() - 1, stream=40 + k.ID) for k in keys] if zkp else None
    plain, item_ok, phases = grp.round(c, rs)
    assert np.array_equal(plain, m) and item_ok.all()
    assert set(phases) == {"pdec", "prove", "all_gather", "verify", "combine"}
    # same plaintexts as the per-share host calls combined by pgpu_combine, and the partial decryptions are libgmp's
    decs = [k.partial_decrypt_records(c) for k in keys]
    assert np.array_equal(decs[0], G.partial_decrypt(n, keys[0].Share, n_dev, c, t0.w_n2))
    assert np.array_equal(t0.combine_records([k.ID for k in keys], np.concatenate(decs)), plain)
    grp.close()
    for k in keys:
        k.close()


def test_library_threshold_round_matches_torchrun_path_on_two_gpus():
    # pgpu_multi_* (single process, ncclCommInitAll) against the one-process-per-GPU path on the same 2-share key; skipped on
    # a single-GPU box
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.path.join(root, "tools", "run_cfg4.py")
    common = ["--bits", "512", "--count", "301", "--shares", "2", "--threshold", "2", "--zkp", "--oracle-items", "64"]
    r1 = subprocess.run([sys.executable, tool, "--impl", "lib"] + common, capture_output=True, text=True, timeout=600, cwd=root)
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-3000:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                         "--master-port", "29534", tool] + common, capture_output=True, text=True, timeout=600, cwd=root)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-3000:]
    for r in (r1, r2):
        line = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
        assert line["n_gpus"] == 2 and line["all_plaintexts_recovered"] and line["oracle_parity"] is True


@pytest.mark.parametrize("bits,l,w,count", [(512, 4, 3, 37), (2048, 8, 5, 9)])
def test_shared_ciphertext_verification_matches_per_server_verification(bits, l, w, count):
    # pgpu_pdec_zkp_verify_shared_dev (the k proofs of one ciphertext share the squarings of (c^4)^Z) against N x VerifyProof per
    # server, honest and tampered proofs alike
    import ctypes as C
    import numpy as np
    from paillier_b200._lib import check, lib
    p, q = synth.load_key(f"threshold_{bits}")
    n = p * q
    keys = ThresholdKeyGenerator(bits, l, w, rng=random.Random(21)).with_safe_primes(p, q).GenerateKeys()
    t0 = keys[0]
    c = t0.encrypt_with_r_records(synth.plaintexts(count, n, t0.w_n), synth.randomness(count, n, t0.w_n))
    c[:t0.w_n2] = 0; c[0] = 1                                                  # c = 1 rides along
    proofs = [k.zkp_prove_records(c, synth.random_records(count, t0.w_n2, (n * n).bit_length() - 1, stream=90 + k.ID)) for k in keys]
    dec = np.stack([pr[0].reshape(count, -1) for pr in proofs], axis=1).copy()   # [ciphertext][share][bytes]
    e = np.stack([pr[1].reshape(count, -1) for pr in proofs], axis=1).copy()
    z = np.stack([pr[2].reshape(count, -1) for pr in proofs], axis=1).copy()
    z[3, 1, 0] ^= 1                       # server 2's Z for ciphertext 3
    e[5, 0, 7] ^= 0x10                    # server 1's E for ciphertext 5
    dec[6, l - 1, 2] ^= 4                 # the last server's partial decryption of ciphertext 6
    dev = torch.device("cuda", 0)
    td = lambda a: torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).to(dev)
    c_d, dec_d, e_d, z_d = td(c), td(dec), td(e), td(z)
    ok = torch.zeros(count * l, dtype=torch.uint8, device=dev)
    ids = (C.c_int * l)(*[k.ID for k in keys])
    vp = lambda t: C.c_void_p(t.data_ptr())
    check(lib.pgpu_pdec_zkp_verify_shared_dev(t0._ctx, count, l, ids, vp(c_d), vp(dec_d), vp(e_d), vp(z_d), vp(ok)), t0._ctx)
    check(lib.pgpu_ctx_sync(t0._ctx), t0._ctx)
    got = ok.cpu().numpy().reshape(count, l)
    for j, k in enumerate(keys):
        want = t0.verify_proof_records(k.ID, c, dec[:, j].copy(), e[:, j].copy(), z[:, j].copy())
        assert np.array_equal(got[:, j], want)
    bad = {(3, 1), (5, 0), (6, l - 1)}
    assert all(bool(got[i, j]) == ((i, j) not in bad) for i in range(count) for j in range(l))
    for k in keys:
        k.close()
