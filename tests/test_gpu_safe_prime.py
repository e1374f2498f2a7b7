"""GPU parity for safe-prime candidate testing (safe_prime.go:147-290): same (p, q, accept) per candidate
byte string as the oracle's restatement of one runGenPrimeRoutine iteration."""
import hashlib
import random

import pytest

from oracle import paillier_ref as R
from paillier_b200 import synth
from paillier_b200.keygen import GenerateSafePrime, miller_rabin, safe_prime_scan

pytestmark = pytest.mark.gpu


def _stream(seed: int, nbytes: int) -> bytes:
    out = bytearray()
    ctr = 0
    while len(out) < nbytes:
        out += hashlib.sha256(f"{seed}:{ctr}".encode()).digest()
        ctr += 1
    return bytes(out[:nbytes])


@pytest.mark.parametrize("p_bits,count", [(16, 3000), (33, 3000), (64, 4000), (127, 2000), (512, 600), (1024, 300), (1536, 60)])
def test_candidate_procedure_matches_oracle(p_bits, count):
    nb = (p_bits - 1 + 7) // 8
    raw = _stream(p_bits, count * nb)
    ps, qs, ok = safe_prime_scan(p_bits, raw)
    n_ok = 0
    for i in range(count):
        op, oq, ook = R.safe_prime_candidate(raw[i * nb:(i + 1) * nb], p_bits)
        assert (qs[i], ok[i]) == (oq, ook), (p_bits, i)
        if op:
            assert ps[i] == op
        n_ok += ook
    if p_bits <= 64:
        assert n_ok > 0            # small sizes: the batch does contain safe primes


def test_known_safe_primes_are_accepted():
    # the golden threshold keys are products of safe primes: feeding (p-1)/2 as the candidate bytes must accept
    for name, bits in (("threshold_512", 256), ("threshold_2048", 1024), ("threshold_3072", 1536)):
        raws = b""
        want = []
        for p in synth.load_key(name):
            assert p.bit_length() == bits
            q = (p - 1) // 2
            raws += q.to_bytes((bits - 1 + 7) // 8, "big")
            want.append((p, q))
        ps, qs, ok = safe_prime_scan(bits, raws)
        assert list(zip(ps, qs)) == want and ok == [True, True]


def test_miller_rabin_known_answers():
    rnd = random.Random(3)
    primes = [2 ** 127 - 1, 2 ** 89 - 1 + 2 ** 126 + 0]      # second one fixed up below
    # 127-bit numbers: a Mersenne prime, Carmichael-like composites, random odds
    cands = [2 ** 127 - 1]
    cands += [(2 ** 63 + 25) * (2 ** 63 + 165) | 1]
    for _ in range(200):
        cands.append(rnd.getrandbits(127) | (1 << 126) | 1)
    got = miller_rabin(127, cands)
    assert got == [R._is_probable_prime(c) for c in cands]
    assert got[0] is True and got[1] is False
    # strong pseudoprimes to base 2 must fall to the later bases: 3215031751 = 151*751*28351 (spsp 2,3,5,7)
    assert miller_rabin(32, [3215031751], rounds=1) == [True]
    assert miller_rabin(32, [3215031751], rounds=4) == [True]
    assert miller_rabin(32, [3215031751], rounds=5) == [False]
    # 1024-bit: the golden safe primes and their neighbours
    p, q = synth.load_key("threshold_2048")
    assert miller_rabin(1024, [p, q, p + 2, q + 2 if (q + 2).bit_length() == 1024 else q - 2]) == [True, True] + [
        R._is_probable_prime(p + 2), R._is_probable_prime(q + 2 if (q + 2).bit_length() == 1024 else q - 2)]


def test_generate_safe_prime_first_in_stream_order():
    # safe_prime_test.go:11-67: shape of the result and the minimum-size error
    state = {"ctr": 0}

    def reader(nbytes):
        state["ctr"] += 1
        return _stream(1000 + state["ctr"], nbytes)

    p, q = GenerateSafePrime(64, reader, batch=4096)
    assert p == 2 * q + 1 and p.bit_length() == 64 and R._is_probable_prime(p) and R._is_probable_prime(q)
    assert (p >> 62) == 3                                     # two most significant bits set (safe_prime.go:58-60)
    # the winner is the FIRST accepted candidate of the stream
    raw = _stream(1001, 4096 * 8)
    first = next((R.safe_prime_candidate(raw[i * 8:(i + 1) * 8], 64) for i in range(4096)
                  if R.safe_prime_candidate(raw[i * 8:(i + 1) * 8], 64)[2]), None)
    if first is not None:
        assert (p, q) == first[:2]
    with pytest.raises(ValueError):
        GenerateSafePrime(5, reader)


def test_keygen_and_threshold_keygen_end_to_end():
    # paillier.go:106-179 and thresholdkey_generator.go:47-55 with the prime searches on the GPU
    from paillier_b200.api import ENC_LEVEL_ONE
    from paillier_b200.keygen import KeyGen, ThresholdKeyGenerator
    rnd = random.Random(12)
    sk, pk = KeyGen(128, rnd)
    assert sk.N.bit_length() in (127, 128) and sk.Lambda % 4 == 0 and sk.K == 1 << 64
    ms = [0, 1, sk.N - 1, 12345]
    cts = pk.EncryptWithRBatch(ms, [3, 5, 7, 11])
    assert sk.DecryptBatch(cts) == ms
    alt = pk.AltEncryptWithRAtLevelBatch(ms, [rnd.randrange(sk.N) for _ in ms], ENC_LEVEL_ONE)
    assert sk.DecryptBatch(alt) == ms
    sk.close(); pk.close()
    with pytest.raises(ValueError):
        KeyGen(63)
    with pytest.raises(ValueError):
        KeyGen(62)
    # threshold keys of a 64-bit n: two 32-bit safe primes found on the GPU (thresholdkey_test.go uses 32-bit keys)
    keys = ThresholdKeyGenerator(64, 4, 3, rng=random.Random(8), batch=2048).GenerateKeys()
    n = keys[0].N
    assert n.bit_length() in (63, 64)
    cs = [c.C for c in keys[0].EncryptWithRBatch([42, 0, n - 1], [5, 7, 9])]
    parts = [k.PartialDecryptBatch(cs) for k in keys[:3]]
    assert keys[0].CombinePartialDecryptionsBatch(parts) == [42, 0, n - 1]
    for k in keys:
        k.close()
