"""Host-side logic of the Python mirror (paillier_b200/api.py) that needs no arithmetic on a batch item and therefore no GPU:
the redraw loop of the randomness-drawing callers (utils.go:36-49), the mapping of proofs that do not fit their records to a
false verdict (thresholdkey.go:278-291 returns false for them), the per-ciphertext outcome of CombinePartialDecryptionsZKP
(thresholdkey.go:164-172), level and argument checks (operations.go:96-140).  The GPU entry points are replaced by stubs that
record what they were given; nothing here computes a ciphertext."""
import math

import numpy as np
import pytest

from paillier_b200 import _lib
from paillier_b200.api import (Ciphertext, ENC_LEVEL_ONE, ENC_LEVEL_TWO, MIXED, PartialDecryptionZKP, PublicKey, ThresholdPublicKey,
                               from_records)

P_, Q_ = 1000003, 1000033


def bare(cls, **attrs):
    """an instance without an engine context (no pgpu_ctx_create): only host-side methods may be called on it"""
    obj = cls.__new__(cls)
    obj._ctx = None
    for k, v in attrs.items():
        setattr(obj, k, v)
    return obj


class ScriptedRand:
    def __init__(self, values):
        self.values = list(values)
        self.calls = 0

    def randrange(self, n):
        self.calls += 1
        return self.values.pop(0) % n if self.values else 1 + self.calls


def test_draw_units_redraws_zero_and_non_units_in_place():
    n = P_ * Q_
    pk = bare(PublicKey, N=n)
    batches = []

    def mod_inverse_batch(xs, modsel):
        batches.append(list(xs))
        assert modsel == _lib.MOD_N and all(x != 0 for x in xs), "zeros are redrawn on the host before the GPU unit test"
        if any(math.gcd(x, n) != 1 for x in xs):
            raise _lib.PgpuError(_lib.PGPU_ERR_NOT_INVERTIBLE, "not invertible")
        return [pow(x, -1, n) for x in xs]

    pk.ModInverseBatch = mod_inverse_batch
    # first draws: a unit, 0, a multiple of p, a unit, a multiple of q; the redraws: 0 again, then units
    rnd = ScriptedRand([5, 0, 3 * P_, 7, 11 * Q_, 0, 13, 17, 19])
    rs = pk._draw_units(5, rnd)
    assert rs == [5, 13, 17, 7, 19]                       # positions 1 (twice), 2 and 4 were redrawn, the others kept
    assert all(math.gcd(r, n) == 1 and 0 < r < n for r in rs)
    assert len(batches) == 2 and batches[0] == [5, 13, 3 * P_, 7, 11 * Q_]    # one GPU unit test per round of draws
    assert pk._draw_units(0, rnd) == [] and len(batches) == 2                 # an empty batch never reaches the GPU
    # another error of the unit test is not swallowed
    def broken(xs, modsel):
        raise _lib.PgpuError(_lib.PGPU_ERR_CUDA, "device lost")
    pk.ModInverseBatch = broken
    with pytest.raises(_lib.PgpuError) as e:
        pk._draw_units(2, ScriptedRand([3, 4]))
    assert e.value.code == _lib.PGPU_ERR_CUDA


def test_default_randomness_is_the_os_csprng():
    n = P_ * Q_
    pk = bare(PublicKey, N=n)
    pk.ModInverseBatch = lambda xs, modsel: xs
    a, b = pk._draw_units(64), pk._draw_units(64)
    assert a != b and all(0 < r < n for r in a + b) and len(set(a + b)) > 120


def test_proofs_that_do_not_fit_their_records_verify_false():
    n = P_ * Q_
    tk = bare(ThresholdPublicKey, N=n, w_n2=8, w_z=12)
    seen = {}

    def verify_proof_records(ID, c, dec, e, z):
        seen.update(ID=ID, c=from_records(c, 8), dec=from_records(dec, 8), e=from_records(e, 32), z=from_records(z, 12))
        return np.ones(len(seen["c"]), dtype=np.uint8)       # the GPU says yes to everything it is shown

    tk.verify_proof_records = verify_proof_records
    n2 = n * n
    good = PartialDecryptionZKP(2, 123, 1 << 255, (1 << 96) - 1, 456)
    proofs = [good,
              PartialDecryptionZKP(2, 123, 1 << 256, 5, 456),          # E is not a SHA-256 digest
              PartialDecryptionZKP(2, 123, 5, 1 << 96, 456),           # Z wider than its record
              PartialDecryptionZKP(2, 123, -1, 5, 456),                # negative values
              PartialDecryptionZKP(2, -123, 5, 5, 456),
              PartialDecryptionZKP(2, n2 + 7, 5, 5, n2 + 9)]           # c, c_i are reduced mod n^2 like the reference's Exp does
    assert tk.VerifyProofBatch(proofs) == [True, False, False, False, False, True]
    assert seen["ID"] == 2
    assert seen["e"] == [1 << 255, 0, 0, 0, 0, 5] and seen["z"] == [(1 << 96) - 1, 0, 0, 0, 0, 5]     # unfit proofs are sent as zeros
    assert seen["c"][5] == 9 and seen["dec"][5] == 7
    with pytest.raises(ValueError):
        tk.VerifyProofBatch([good, PartialDecryptionZKP(3, 1, 1, 1, 1)])      # one server per batch
    assert tk.VerifyProofBatch([]) == []


def test_combine_zkp_reports_per_ciphertext():
    n = P_ * Q_
    tk = bare(ThresholdPublicKey, N=n, w_n=4, w_n2=8, w_z=12)
    verdicts = {1: [True, False, True], 2: [True, True, True], 3: [True, False, False]}
    tk.VerifyProofBatch = lambda proofs: verdicts[proofs[0].ID]
    got = {}

    def combine_verified_records(ids, decs, ok):
        got.update(ids=list(ids), ok=list(ok), decs=from_records(decs, 8))
        item_ok = np.array([1, 0, 1], dtype=np.uint8)        # ciphertext 1 is left with one share of three: below the threshold
        out = np.zeros(3 * 4, dtype=np.uint8)
        out[0], out[8] = 42, 44
        return out, item_ok

    tk.combine_verified_records = combine_verified_records
    shares = [[PartialDecryptionZKP(j, 100 * j + i, 1, 1, 7 + i) for i in range(3)] for j in (1, 2, 3)]
    assert tk.CombinePartialDecryptionsZKPBatch(shares, strict=False) == [42, None, 44]
    assert got["ids"] == [1, 2, 3] and got["ok"] == [1, 0, 1, 1, 1, 1, 1, 0, 0]            # server-major, like the partial decryptions
    assert got["decs"] == [100, 101, 102, 200, 201, 202, 300, 301, 302]
    with pytest.raises(_lib.PgpuError) as e:
        tk.CombinePartialDecryptionsZKPBatch(shares)         # strict: any ciphertext below the threshold is the reference's error
    assert e.value.code == _lib.PGPU_ERR_THRESHOLD and "1 of 3" in str(e.value)
    with pytest.raises(_lib.PgpuError):
        tk.CombinePartialDecryptionsZKPBatch([])
    with pytest.raises(_lib.PgpuError):                      # thresholdkey.go:80-88: two shares of the same server
        tk.CombinePartialDecryptionsZKPBatch([shares[0], shares[0]])
    with pytest.raises(ValueError):
        tk.CombinePartialDecryptionsZKPBatch([shares[0], shares[1][:2]])
    # VerifyDecryption compares the ciphertexts first (thresholdkey.go:176-181)
    with pytest.raises(ValueError, match="encrypted message"):
        tk.VerifyDecryptionBatch([7, 8, 10], [42, 43, 44], shares)


def test_level_and_argument_checks_need_no_gpu():
    n = P_ * Q_
    pk = bare(PublicKey, N=n, H=None, K=None, w_n=4, w_n2=8, w_n3=12)
    one, two = Ciphertext(5, ENC_LEVEL_ONE), Ciphertext(6, ENC_LEVEL_TWO)
    with pytest.raises(ValueError, match="one encryption level per batch"):
        pk.AddPairs([one, two], [one, one])
    with pytest.raises(ValueError, match="one encryption level per batch"):
        pk.ConstMultBatch([one, two], [1, 2])
    with pytest.raises(ValueError, match="doubly encrypted"):
        pk.NestedRandomizeWithBatch([one], [3], [4])                      # operations.go:97-99 panics
    with pytest.raises(ValueError, match="doubly encrypted"):
        pk.NestedAddBatch([one], [one])                                    # operations.go:122-124
    with pytest.raises(ValueError, match="doubly encrypted"):
        pk.NestedSubBatch([two], [two])
    with pytest.raises(ValueError, match="H and K"):
        pk.AltEncryptWithRAtLevelBatch([1], [2], ENC_LEVEL_ONE)
    with pytest.raises(ValueError):
        pk.EncryptWithRBatch([1, 2], [3])
    with pytest.raises(ValueError):
        pk.SubBatch([])
    big = Ciphertext(n ** 2 + 5, ENC_LEVEL_ONE)
    only = pk.SubBatch([big])                                              # operations.go:34: one argument comes back unreduced
    assert (only.C, only.Level, only.EncMethod) == (n ** 2 + 5, ENC_LEVEL_ONE, MIXED)
    assert pk.GetN2() == n * n and pk.GetN3() == n ** 3
    pk3 = bare(PublicKey, N=n, w_n=4, w_n2=8, w_n3=0)
    with pytest.raises(_lib.PgpuError) as e:
        pk3._level_modulus(ENC_LEVEL_TWO)                                  # n^3 wider than the built kernel shapes
    assert e.value.code == _lib.PGPU_ERR_UNSUPPORTED
    assert pk._one_level([], "x") == ENC_LEVEL_ONE and pk._one_level([two, two], "x") == ENC_LEVEL_TWO
