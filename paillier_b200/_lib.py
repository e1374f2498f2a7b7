"""ctypes binding of libpaillier_b200.so (C ABI declared in include/pgpu.h).

There is no fallback: if the shared library has not been built the import
fails, and without a CUDA device every compute entry point returns
PGPU_ERR_CUDA, which `check` turns into an exception.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGPU_LIB_PATH") or os.path.join(_HERE, "libpaillier_b200.so")

PGPU_OK, PGPU_ERR_ARG, PGPU_ERR_CUDA, PGPU_ERR_NCCL, PGPU_ERR_STATE = 0, 1, 2, 3, 4
PGPU_ERR_NOT_INVERTIBLE, PGPU_ERR_THRESHOLD, PGPU_ERR_UNSUPPORTED = 5, 6, 7
MOD_N, MOD_N2, MOD_N3 = 0, 1, 2

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C paillier_b200/csrc` (there is no CPU fallback)")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_sz = C.c_size_t
_u8p = C.c_char_p

# name -> (restype, argtypes); every symbol include/pgpu.h declares
SIGNATURES = {
    "pgpu_version": (C.c_int, []),
    "pgpu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pgpu_last_error": (C.c_char_p, [_p]),
    "pgpu_ctx_create": (C.c_int, [C.POINTER(_p), C.c_int, _u8p, _sz]),
    "pgpu_ctx_destroy": (C.c_int, [_p]),
    "pgpu_ctx_widths": (C.c_int, [_p, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "pgpu_ctx_mod_width": (C.c_int, [_p, C.c_int, C.POINTER(_sz)]),
    "pgpu_ctx_set_stream": (C.c_int, [_p, _p]),
    "pgpu_ctx_set_secret_pq": (C.c_int, [_p, _u8p, _sz, _u8p, _sz]),
    "pgpu_ctx_set_secret_lambda": (C.c_int, [_p, _u8p, _sz]),
    "pgpu_ctx_set_threshold": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _u8p, _sz, _u8p, _sz, _p]),
    "pgpu_encrypt_with_r": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_decrypt": (C.c_int, [_p, _sz, _p, _p]),
    "pgpu_const_mult": (C.c_int, [_p, _sz, _p, _p, _sz, _p]),
    "pgpu_add_reduce": (C.c_int, [_p, _sz, _p, _p]),
    "pgpu_add_reduce_at_level": (C.c_int, [_p, C.c_int, _sz, _p, _p]),
    "pgpu_add_pairs": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_dot_u64": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_partial_decrypt": (C.c_int, [_p, _sz, _p, _p]),
    "pgpu_modexp": (C.c_int, [_p, C.c_int, _sz, _p, _p, _sz, _p]),
    "pgpu_modexp_shared": (C.c_int, [_p, C.c_int, _sz, _p, _u8p, _sz, _p]),
    "pgpu_modmul": (C.c_int, [_p, C.c_int, _sz, _p, _p, _p]),
    "pgpu_buf_alloc": (C.c_int, [_p, _sz, C.POINTER(_p)]),
    "pgpu_buf_free": (C.c_int, [_p]),
    "pgpu_buf_ptr": (C.c_void_p, [_p]),
    "pgpu_buf_size": (C.c_size_t, [_p]),
    "pgpu_buf_upload": (C.c_int, [_p, _sz, _p, _sz]),
    "pgpu_buf_download": (C.c_int, [_p, _sz, _p, _sz]),
    "pgpu_ctx_sync": (C.c_int, [_p]),
    "pgpu_host_alloc": (C.c_int, [_sz, C.POINTER(_p)]),
    "pgpu_host_free": (C.c_int, [_p]),
    "pgpu_encrypt_with_r_dev": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_encrypt_with_r_sk": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_encrypt_with_rn": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_encrypt_with_rn_dev": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_encrypt_with_r_sk_dev": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_decrypt_dev": (C.c_int, [_p, _sz, _p, _p]),
    "pgpu_partial_decrypt_dev": (C.c_int, [_p, _sz, _p, _p]),
    "pgpu_const_mult_dev": (C.c_int, [_p, _sz, _p, _p, _sz, _p]),
    "pgpu_add_reduce_dev": (C.c_int, [_p, _sz, _p, _p]),
    "pgpu_dot_u64_dev": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_add_pairs_dev": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_sub_pairs_dev": (C.c_int, [_p, _sz, _p, _p, _p, _p]),
    "pgpu_randomize_with_r_dev": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_sub_pairs": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_modinv": (C.c_int, [_p, C.c_int, _sz, _p, _p]),
    "pgpu_pdec_zkp_prove": (C.c_int, [_p, _sz, _p, _p, _p, _p, _p]),
    "pgpu_ctx_z_width": (C.c_int, [_p, C.POINTER(_sz)]),
    "pgpu_pdec_zkp_verify": (C.c_int, [_p, _sz, C.c_int, _p, _p, _p, _p, _p]),
    "pgpu_combine": (C.c_int, [_p, _sz, C.c_int, C.POINTER(C.c_int), _p, _p]),
    "pgpu_pdec_zkp_verify_shared_dev": (C.c_int, [_p, _sz, C.c_int, C.POINTER(C.c_int), _p, _p, _p, _p, _p]),
    "pgpu_combine_verified": (C.c_int, [_p, _sz, C.c_int, C.POINTER(C.c_int), _p, _p, _p, _p]),
    "pgpu_combine_verified_dev": (C.c_int, [_p, _sz, C.c_int, C.POINTER(C.c_int), _p, _sz, _p, _p, _p]),
    "pgpu_pdec_zkp_prove_dev": (C.c_int, [_p, _sz, _p, _p, _p, _p, _p]),
    "pgpu_pdec_zkp_prove_given_dev": (C.c_int, [_p, _sz, _p, _p, _p, _p, _p]),
    "pgpu_combine_dev": (C.c_int, [_p, _sz, C.c_int, C.POINTER(C.c_int), _p, _p]),
    "pgpu_ctx_set_alt_generator": (C.c_int, [_p, _u8p, _sz, C.c_uint]),
    "pgpu_encrypt_with_r_at_level": (C.c_int, [_p, C.c_int, _sz, _p, _p, _p]),
    "pgpu_encrypt_with_r_at_level_sk": (C.c_int, [_p, C.c_int, _sz, _p, _p, _p]),
    "pgpu_alt_encrypt_with_r_at_level": (C.c_int, [_p, C.c_int, _sz, _p, _p, _p]),
    "pgpu_decrypt_at_level": (C.c_int, [_p, C.c_int, _sz, _p, _p]),
    "pgpu_randomize_with_r": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_extract_randomness": (C.c_int, [_p, C.c_int, _sz, _p, _p]),
    "pgpu_nested_randomize_with": (C.c_int, [_p, _sz, _p, _p, _p, _p]),
    "pgpu_nested_add": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_nested_sub": (C.c_int, [_p, _sz, _p, _p, _p]),
    "pgpu_ddleq_prove": (C.c_int, [_p, _sz, C.c_uint, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pgpu_ddleq_verify": (C.c_int, [_p, _sz, C.c_uint, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pgpu_safe_prime_scan": (C.c_int, [C.c_int, C.c_uint, _sz, _p, _p, _p, _p, C.POINTER(C.c_uint64)]),
    "pgpu_miller_rabin": (C.c_int, [C.c_int, C.c_uint, _sz, _p, C.c_uint, _p, C.POINTER(C.c_uint64)]),
    "pgpu_primes_last_error": (C.c_char_p, []),
    "pgpu_combine_strided_dev": (C.c_int, [_p, _sz, C.c_int, C.POINTER(C.c_int), _p, _sz, _p]),
    "pgpu_pdec_zkp_verify_dev": (C.c_int, [_p, _sz, C.c_int, _p, _p, _p, _p, _p]),
    "pgpu_pdec_zkp_verify_multi_dev": (C.c_int, [_p, _sz, C.c_int, C.POINTER(C.c_int), _p, _p, _p, _p, _p]),
    "pgpu_multi_create": (C.c_int, [C.POINTER(_p), C.POINTER(_p), C.c_int]),
    "pgpu_multi_destroy": (C.c_int, [_p]),
    "pgpu_multi_size": (C.c_int, [_p]),
    "pgpu_multi_last_error": (C.c_char_p, [_p]),
    "pgpu_multi_threshold_round": (C.c_int, [_p, _sz, _p, C.POINTER(_p), _p, _p]),
    "pgpu_multi_last_phases_ms": (C.c_int, [_p, C.POINTER(C.c_float)]),
    "pgpu_ctx_launch_count": (C.c_int, [_p, C.POINTER(C.c_uint64)]),
    "pgpu_ctx_program_cost": (C.c_int, [_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "pgpu_ctx_kernel_shape": (C.c_int, [_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pgpu_ctx_enable_timing": (C.c_int, [_p, C.c_int]),
    "pgpu_ctx_last_kernel_ms": (C.c_int, [_p, C.POINTER(C.c_float)]),
    "pgpu_selftest_program": (C.c_int, [C.c_int, _u8p, _sz, _u8p, _sz, _u8p, _sz, _p, C.c_uint32, C.c_uint32, C.c_uint32, _p, _sz,
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "pgpu_selftest_bn": (C.c_int, [C.c_int, _u8p, _sz, _u8p, _sz, _u8p, _sz, C.c_char_p, C.POINTER(_sz)]),
    "pgpu_selftest_last_program": (C.c_int, [_p, _sz, C.POINTER(_sz), C.POINTER(C.c_uint32)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


class PgpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pgpu error {code}: {msg}")
        self.code = code


def check(rc: int, ctx=None) -> None:
    if rc != PGPU_OK:
        msg = lib.pgpu_last_error(ctx) or b""
        raise PgpuError(rc, msg.decode("utf-8", "replace"))
