"""The FP64-pipe exponentiation kernels (csrc/mont52.cuh, powm_vm52) at every built width.

By default only 4096-bit moduli run on them (csrc/engine.cu: pick_shape); here the parity suites that compare the C-ABI
with the oracles and the golden vectors are re-run in a child process with every width switched to its FP64 shape
(PGPU_SHAPE_<S>="tpi,L,fp64"), and once more with every width on the integer pipe (PGPU_NO_FP64=1), so both multipliers
stay bit-exact at 1024 ... 6144 bits whichever one a width is served by."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITES = ["tests/test_gpu_golden.py", "tests/test_gpu_level2_ddleq.py", "tests/test_gpu_parity.py"]


def _run(env_extra, select=None, suites=SUITES):
    env = dict(os.environ)
    env.update(env_extra)
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + list(suites)
    if select:
        cmd += ["-k", select]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return r.stdout


def test_every_width_on_the_fp64_pipe():
    out = _run({"PGPU_SHAPE_32": "4,5,fp64", "PGPU_SHAPE_64": "4,10,fp64", "PGPU_SHAPE_96": "4,15,fp64",
                "PGPU_SHAPE_128": "8,10,fp64", "PGPU_SHAPE_192": "8,15,fp64"},
               select="golden or level2_encrypt or (ddleq_prove and not paillier_2048) or encrypt_decrypt_parity or partial_decrypt_parity or zkp_prove_verify "
                      "or carry_chain or sub_and")
    assert " passed" in out


def test_alternate_fp64_shapes():
    out = _run({"PGPU_SHAPE_64": "8,5,fp64", "PGPU_SHAPE_96": "8,8,fp64", "PGPU_SHAPE_192": "16,8,fp64"}, suites=["tests/test_gpu_golden.py"])
    assert " passed" in out


def test_every_width_on_the_integer_pipe():
    out = _run({"PGPU_NO_FP64": "1"}, select="golden or encrypt_decrypt_parity or partial_decrypt_parity or carry_chain")
    assert " passed" in out
