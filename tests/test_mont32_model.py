"""CPU model of the integer-pipe Montgomery multiplication of paillier_b200/csrc/mont.cuh (no GPU needed).

Mont<TPI, L>::mul keeps a lane's accumulator as an even and an odd array of 64-bit columns that swap roles at every row, with
the one-limb shift folded into the addends of the next row, lazy per-lane carry words, and one ballot-based carry look-ahead plus
conditional subtraction at the end (resolve_reduce).  tools/sim_mont.py runs exactly that bookkeeping lane by lane on Python
integers -- the column pairing, which carry feeds which chain, what crosses to the neighbouring lane, the bounds of the carry
words -- for every shape powm.cu builds, and checks r = a*b*R^-1 mod n (canonical, < n) for random and extreme moduli and
operands.  The instruction-level arithmetic itself (mad.lo.cc / madc.hi.cc chains) is exact by construction; the GPU parity
tests cover it and the dedicated squaring."""
import importlib.util
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("sim_mont", os.path.join(ROOT, "tools", "sim_mont.py"))
sim = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sim)


def built_shapes():
    src = open(os.path.join(ROOT, "paillier_b200", "csrc", "powm.cu")).read()
    body = src[src.index("#define PGPU_FOR_EACH_SHAPE(X)"):src.index("// FP64-pipe shapes")]
    return [(int(a), int(b)) for a, b in re.findall(r"X\((\d+),\s*(\d+)\)", body)]


SHAPES = built_shapes()


def test_shape_list_is_the_built_one():
    assert len(SHAPES) >= 13 and (4, 32) in SHAPES and (8, 24) in SHAPES and (4, 16) in SHAPES
    assert all(l % 2 == 0 and t & (t - 1) == 0 for t, l in SHAPES)


@pytest.mark.parametrize("tpi,L", SHAPES)
def test_lane_sliced_cios_matches_big_integers(tpi, L):
    # moduli: full width, R - small, short, 2^(32S-1) + small; operands: random, n-1, R-1, 0, 1 (tools/sim_mont.py: test)
    sim.test(tpi, L, 12 if tpi * L > 64 else 40)
