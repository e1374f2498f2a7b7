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
