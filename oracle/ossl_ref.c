/*
 * ossl_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A third bignum implementation, unrelated to CPython ints and to libgmp, for the large-size golden vectors
 * (SURVEY.md 8c "independent cross-checks": OpenSSL BN): gmp.Int.Exp / Mul+Mod / ModInverse answered by
 * BN_mod_exp / BN_mod_mul / BN_mod_inverse over the same fixed-width little-endian records.
 * tests/test_golden_oracle.py re-derives tests/golden/vectors.json with the Python oracle's primitives bound to it.
 *
 * Build: see oracle/Makefile  (gcc -O2 -shared -fPIC ossl_ref.c -l:libcrypto.so.3)
 */
#include <openssl/bn.h>
#include <stdint.h>
#include <string.h>

/* ncw/gmp Int.Exp semantics are applied by the caller (y <= 0 -> 1, nil modulus -> plain power) */
int ossl_modexp(const uint8_t* mod_be, size_t mod_len, size_t count, const uint8_t* base, size_t width,
                const uint8_t* exp, size_t exp_bytes, uint8_t* out) {
    BN_CTX* ctx = BN_CTX_new();
    BIGNUM *m = BN_bin2bn(mod_be, (int)mod_len, NULL), *b = BN_new(), *e = BN_new(), *r = BN_new();
    int rc = 0;
    for (size_t i = 0; i < count && rc == 0; ++i) {
        BN_lebin2bn(base + i * width, (int)width, b);
        BN_lebin2bn(exp + i * exp_bytes, (int)exp_bytes, e);
        if (!BN_nnmod(b, b, m, ctx) || !BN_mod_exp(r, b, e, m, ctx)) rc = 1;
        else if (BN_bn2lebinpad(r, out + i * width, (int)width) < 0) rc = 2;
    }
    BN_free(m); BN_free(b); BN_free(e); BN_free(r); BN_CTX_free(ctx);
    return rc;
}

int ossl_modmul(const uint8_t* mod_be, size_t mod_len, size_t count, const uint8_t* a, const uint8_t* b, size_t width, uint8_t* out) {
    BN_CTX* ctx = BN_CTX_new();
    BIGNUM *m = BN_bin2bn(mod_be, (int)mod_len, NULL), *x = BN_new(), *y = BN_new(), *r = BN_new();
    int rc = 0;
    for (size_t i = 0; i < count && rc == 0; ++i) {
        BN_lebin2bn(a + i * width, (int)width, x);
        BN_lebin2bn(b + i * width, (int)width, y);
        if (!BN_mod_mul(r, x, y, m, ctx)) rc = 1;
        else if (BN_bn2lebinpad(r, out + i * width, (int)width) < 0) rc = 2;
    }
    BN_free(m); BN_free(x); BN_free(y); BN_free(r); BN_CTX_free(ctx);
    return rc;
}

/* ok[i] = 0 where a[i] has no inverse (record left zero) */
int ossl_modinv(const uint8_t* mod_be, size_t mod_len, size_t count, const uint8_t* a, size_t width, uint8_t* out, uint8_t* ok) {
    BN_CTX* ctx = BN_CTX_new();
    BIGNUM *m = BN_bin2bn(mod_be, (int)mod_len, NULL), *x = BN_new(), *r = BN_new();
    for (size_t i = 0; i < count; ++i) {
        BN_lebin2bn(a + i * width, (int)width, x);
        memset(out + i * width, 0, width);
        ok[i] = 0;
        if (BN_mod_inverse(r, x, m, ctx)) { ok[i] = 1; BN_bn2lebinpad(r, out + i * width, (int)width); }
    }
    BN_free(m); BN_free(x); BN_free(r); BN_CTX_free(ctx);
    return 0;
}
