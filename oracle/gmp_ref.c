/*
 * gmp_ref.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (never linked into the product).
 *
 * The reference does all hot-path arithmetic through github.com/ncw/gmp, a cgo
 * wrapper over libgmp's mpz_* (SURVEY.md fact 1).  Go and ncw/gmp are not
 * available here, so this file restates, function by function, the exact
 * sequence of libgmp calls the reference issues -- no g=n+1 shortcut, no CRT,
 * no precomputation the reference does not do -- over the same fixed-width
 * little-endian records the GPU library uses.  It serves as
 *   (1) the second, independent oracle for large sizes (validated against
 *       oracle/paillier_ref.py in tests/test_oracle_gmp.py), and
 *   (2) bench.py's cpu_baseline / `--impl reference` arm ("port": a libgmp
 *       stand-in for the Go package; it omits cgo and GC overhead, so it is
 *       faster than the real thing would be).
 *
 * libgmp.so.10 (GMP 6.3.0) is in the image without headers, hence the
 * hand-declared prototypes below.
 *
 * Build: see oracle/Makefile  (gcc -O2 -shared -fPIC gmp_ref.c -l:libgmp.so.10 -lpthread)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long mp_limb_t;
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t* _mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];

extern void __gmpz_init(mpz_t);
extern void __gmpz_clear(mpz_t);
extern void __gmpz_set(mpz_t, const mpz_t);
extern void __gmpz_set_ui(mpz_t, unsigned long);
extern void __gmpz_import(mpz_t, size_t, int, size_t, int, size_t, const void*);
extern void* __gmpz_export(void*, size_t*, int, size_t, int, size_t, const mpz_t);
extern void __gmpz_powm(mpz_t, const mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_pow_ui(mpz_t, const mpz_t, unsigned long);
extern void __gmpz_mul(mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_mul_ui(mpz_t, const mpz_t, unsigned long);
extern void __gmpz_mod(mpz_t, const mpz_t, const mpz_t);
extern int __gmpz_invert(mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_add(mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_add_ui(mpz_t, const mpz_t, unsigned long);
extern void __gmpz_sub(mpz_t, const mpz_t, const mpz_t);
extern void __gmpz_sub_ui(mpz_t, const mpz_t, unsigned long);
extern void __gmpz_fdiv_q(mpz_t, const mpz_t, const mpz_t);
extern int __gmpz_cmp_ui(const mpz_t, unsigned long);
extern int __gmpz_cmp(const mpz_t, const mpz_t);

#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_set __gmpz_set
#define mpz_set_ui __gmpz_set_ui
#define mpz_powm __gmpz_powm
#define mpz_pow_ui __gmpz_pow_ui
#define mpz_mul __gmpz_mul
#define mpz_mul_ui __gmpz_mul_ui
#define mpz_mod __gmpz_mod
#define mpz_invert __gmpz_invert
#define mpz_add __gmpz_add
#define mpz_add_ui __gmpz_add_ui
#define mpz_sub __gmpz_sub
#define mpz_sub_ui __gmpz_sub_ui
#define mpz_fdiv_q __gmpz_fdiv_q
#define mpz_cmp_ui __gmpz_cmp_ui
#define mpz_cmp __gmpz_cmp
#define mpz_sgn(z) ((z)->_mp_size < 0 ? -1 : (z)->_mp_size > 0)

static void imp_be(mpz_t z, const uint8_t* p, size_t len) { __gmpz_import(z, len, 1, 1, 1, 0, p); }
static void imp_le(mpz_t z, const uint8_t* p, size_t len) { __gmpz_import(z, len, -1, 1, -1, 0, p); }
static void exp_le(uint8_t* out, size_t width, const mpz_t z) {
    size_t cnt = 0;
    memset(out, 0, width);
    uint8_t* tmp = (uint8_t*)__gmpz_export(NULL, &cnt, -1, 1, -1, 0, z);
    if (tmp) { memcpy(out, tmp, cnt < width ? cnt : width); free(tmp); }
}

/* ncw/gmp Int.Exp: y <= 0 -> 1; m == nil / 0 -> mpz_pow_ui; else mpz_powm */
static void gmp_exp(mpz_t z, const mpz_t x, const mpz_t y, const mpz_t m) {
    if (mpz_sgn(y) <= 0) { mpz_set_ui(z, 1); return; }
    mpz_powm(z, x, y, m);
}

typedef struct job job_t;
typedef void (*item_fn)(const job_t*, size_t lo, size_t hi);
struct job {
    item_fn fn;
    size_t lo, hi;
    /* shared, read-only */
    const uint8_t *a, *b; uint8_t* out;
    size_t wa, wb, wout;
    mpz_t k0, k1, k2, k3;
    unsigned long u0;
};

static void* trampoline(void* p) { const job_t* j = (const job_t*)p; j->fn(j, j->lo, j->hi); return NULL; }

static void run_parallel(job_t* proto, size_t count, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > count) threads = count ? (int)count : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = *proto;
        jobs[t].lo = count * t / threads;
        jobs[t].hi = count * (t + 1) / threads;
        pthread_create(&th[t], NULL, trampoline, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ---- EncryptWithRAtLevel, level 1 (paillier.go:206-218): k0 = n, k1 = n^2, k2 = g ---- */
static void enc_items(const job_t* j, size_t lo, size_t hi) {
    mpz_t m, r, gm, rn, c;
    mpz_init(m); mpz_init(r); mpz_init(gm); mpz_init(rn); mpz_init(c);
    for (size_t i = lo; i < hi; ++i) {
        imp_le(m, j->a + i * j->wa, j->wa);
        imp_le(r, j->b + i * j->wb, j->wb);
        gmp_exp(gm, j->k2, m, j->k1);          /* gm := Exp(pk.G, m, ns1)   :213 */
        gmp_exp(rn, r, j->k0, j->k1);          /* rn := Exp(r, ns, ns1)     :214 */
        mpz_mul(c, gm, rn);                    /* c := Mod(Mul(gm, rn), ns1) :216 */
        mpz_mod(c, c, j->k1);
        exp_le(j->out + i * j->wout, j->wout, c);
    }
    mpz_clear(m); mpz_clear(r); mpz_clear(gm); mpz_clear(rn); mpz_clear(c);
}

int ref_encrypt_with_r(const uint8_t* n_be, size_t n_len, size_t count, const uint8_t* m, const uint8_t* r,
                       size_t w_n, uint8_t* out, size_t w_n2, int threads) {
    job_t j; memset(&j, 0, sizeof j);
    mpz_init(j.k0); mpz_init(j.k1); mpz_init(j.k2);
    imp_be(j.k0, n_be, n_len);
    mpz_mul(j.k1, j.k0, j.k0);                 /* GetN2, paillier.go:72-79 */
    mpz_add_ui(j.k2, j.k0, 1);                 /* g = n + 1, paillier.go:147 */
    j.fn = enc_items; j.a = m; j.b = r; j.out = out; j.wa = w_n; j.wb = w_n; j.wout = w_n2;
    run_parallel(&j, count, threads);
    mpz_clear(j.k0); mpz_clear(j.k1); mpz_clear(j.k2);
    return 0;
}

/* ---- Decrypt, level 1 (paillier.go:292-303 + recoveryAlgorithm :308-340 + L :437-440):
 *      k0 = n, k1 = n^2, k2 = lambda ---- */
static void dec_items(const job_t* j, size_t lo, size_t hi) {
    mpz_t c, tmp, nj, nj1, amod, t1, mu, mm;
    mpz_init(c); mpz_init(tmp); mpz_init(nj); mpz_init(nj1); mpz_init(amod); mpz_init(t1); mpz_init(mu); mpz_init(mm);
    for (size_t i = lo; i < hi; ++i) {
        imp_le(c, j->a + i * j->wa, j->wa);
        gmp_exp(tmp, c, j->k2, j->k1);         /* tmp := Exp(ct.C, sk.Lambda, ns1)  :296 */
        mpz_pow_ui(nj, j->k0, 1);              /* nj := Exp(N, j, nil)              :313 */
        mpz_pow_ui(nj1, j->k0, 2);             /* nj1 := Exp(N, j+1, nil)           :314 */
        mpz_mod(amod, tmp, nj1);               /* :316 */
        mpz_sub_ui(t1, amod, 1);               /* L: (u-1)/n                        :437-440 */
        mpz_fdiv_q(t1, t1, j->k0);
        if (!mpz_invert(mu, j->k2, j->k0)) mpz_set_ui(mu, 0);   /* mu := ModInverse(Lambda, ns) :298 */
        mpz_mul(mm, t1, mu);                   /* m := Mod(Mul(ml, mu), ns)         :300 */
        mpz_mod(mm, mm, j->k0);
        exp_le(j->out + i * j->wout, j->wout, mm);
    }
    mpz_clear(c); mpz_clear(tmp); mpz_clear(nj); mpz_clear(nj1); mpz_clear(amod); mpz_clear(t1); mpz_clear(mu); mpz_clear(mm);
}

int ref_decrypt(const uint8_t* n_be, size_t n_len, const uint8_t* lambda_be, size_t lambda_len, size_t count,
                const uint8_t* c, size_t w_n2, uint8_t* out, size_t w_n, int threads) {
    job_t j; memset(&j, 0, sizeof j);
    mpz_init(j.k0); mpz_init(j.k1); mpz_init(j.k2);
    imp_be(j.k0, n_be, n_len);
    mpz_mul(j.k1, j.k0, j.k0);
    imp_be(j.k2, lambda_be, lambda_len);
    j.fn = dec_items; j.a = c; j.out = out; j.wa = w_n2; j.wout = w_n;
    run_parallel(&j, count, threads);
    mpz_clear(j.k0); mpz_clear(j.k1); mpz_clear(j.k2);
    return 0;
}

/* ---- PartialDecrypt (thresholdkey.go:192-201): k0 = n, k2 = share, u0 = l ---- */
static void pdec_items(const job_t* j, size_t lo, size_t hi) {
    mpz_t c, delta, e, n2, r;
    mpz_init(c); mpz_init(delta); mpz_init(e); mpz_init(n2); mpz_init(r);
    for (size_t i = lo; i < hi; ++i) {
        imp_le(c, j->a + i * j->wa, j->wa);
        mpz_set_ui(delta, 1);                  /* delta() = Factorial(l), recomputed per call :70-72, utils.go:17-23 */
        for (unsigned long k = 1; k <= j->u0; ++k) mpz_mul_ui(delta, delta, k);
        mpz_mul_ui(e, delta, 2);               /* exp := Share * (2 * delta)   :195 */
        mpz_mul(e, j->k2, e);
        mpz_mul(n2, j->k0, j->k0);             /* GetN2() */
        gmp_exp(r, c, e, n2);                  /* Exp(gmpC, gmpExp, gmpN2)     :199 */
        exp_le(j->out + i * j->wout, j->wout, r);
    }
    mpz_clear(c); mpz_clear(delta); mpz_clear(e); mpz_clear(n2); mpz_clear(r);
}

int ref_partial_decrypt(const uint8_t* n_be, size_t n_len, const uint8_t* share_be, size_t share_len, int l,
                        size_t count, const uint8_t* c, uint8_t* out, size_t w_n2, int threads) {
    job_t j; memset(&j, 0, sizeof j);
    mpz_init(j.k0); mpz_init(j.k2);
    imp_be(j.k0, n_be, n_len);
    imp_be(j.k2, share_be, share_len);
    j.u0 = (unsigned long)l;
    j.fn = pdec_items; j.a = c; j.out = out; j.wa = w_n2; j.wout = w_n2;
    run_parallel(&j, count, threads);
    mpz_clear(j.k0); mpz_clear(j.k2);
    return 0;
}

/* ---- generic Exp / Mul+Mod against one modulus k1 (ConstMult operations.go:58-64,
 *      Add operations.go:17-22, and the ZKP / DDLEQ call sites) ---- */
static void modexp_items(const job_t* j, size_t lo, size_t hi) {
    mpz_t b, e, r;
    mpz_init(b); mpz_init(e); mpz_init(r);
    for (size_t i = lo; i < hi; ++i) {
        imp_le(b, j->a + i * j->wa, j->wa);
        imp_le(e, j->b + i * j->wb, j->wb);
        gmp_exp(r, b, e, j->k1);
        exp_le(j->out + i * j->wout, j->wout, r);
    }
    mpz_clear(b); mpz_clear(e); mpz_clear(r);
}

int ref_modexp(const uint8_t* mod_be, size_t mod_len, size_t count, const uint8_t* base, size_t width,
               const uint8_t* exp, size_t exp_bytes, uint8_t* out, int threads) {
    job_t j; memset(&j, 0, sizeof j);
    mpz_init(j.k1);
    imp_be(j.k1, mod_be, mod_len);
    j.fn = modexp_items; j.a = base; j.b = exp; j.out = out; j.wa = width; j.wb = exp_bytes; j.wout = width;
    run_parallel(&j, count, threads);
    mpz_clear(j.k1);
    return 0;
}

static void modmul_items(const job_t* j, size_t lo, size_t hi) {
    mpz_t a, b;
    mpz_init(a); mpz_init(b);
    for (size_t i = lo; i < hi; ++i) {
        imp_le(a, j->a + i * j->wa, j->wa);
        imp_le(b, j->b + i * j->wb, j->wb);
        mpz_mul(a, a, b);
        mpz_mod(a, a, j->k1);
        exp_le(j->out + i * j->wout, j->wout, a);
    }
    mpz_clear(a); mpz_clear(b);
}

int ref_modmul(const uint8_t* mod_be, size_t mod_len, size_t count, const uint8_t* a, const uint8_t* b,
               size_t width, uint8_t* out, int threads) {
    job_t j; memset(&j, 0, sizeof j);
    mpz_init(j.k1);
    imp_be(j.k1, mod_be, mod_len);
    j.fn = modmul_items; j.a = a; j.b = b; j.out = out; j.wa = width; j.wb = width; j.wout = width;
    run_parallel(&j, count, threads);
    mpz_clear(j.k1);
    return 0;
}

/* ---- gmp.Int.ModInverse (mpz_invert) per item: Sub (operations.go:46-50), verifyPart1/2 (thresholdkey.go:293-311),
 *      proveDDLEQInstance (ddleq.go:96,110).  ok[i] = 0 where no inverse exists (the record is left zero). ---- */
static void modinv_items(const job_t* j, size_t lo, size_t hi) {
    mpz_t a;
    mpz_init(a);
    for (size_t i = lo; i < hi; ++i) {
        imp_le(a, j->a + i * j->wa, j->wa);
        const int ok = mpz_invert(a, a, j->k1);
        ((uint8_t*)j->b)[i] = (uint8_t)(ok != 0);
        if (ok) exp_le(j->out + i * j->wout, j->wout, a);
    }
    mpz_clear(a);
}

int ref_modinv(const uint8_t* mod_be, size_t mod_len, size_t count, const uint8_t* a, size_t width, uint8_t* out,
               uint8_t* ok, int threads) {
    job_t j; memset(&j, 0, sizeof j);
    mpz_init(j.k1);
    imp_be(j.k1, mod_be, mod_len);
    j.fn = modinv_items; j.a = a; j.b = ok; j.out = out; j.wa = width; j.wout = width;
    run_parallel(&j, count, threads);
    mpz_clear(j.k1);
    return 0;
}

/* ---- Add over a batch (operations.go:11-29): accumulator := Mod(Mul(accumulator, c), ns1).
 *      Each thread folds its slice; partial products are folded at the end (same value). ---- */
static void fold_items(const job_t* j, size_t lo, size_t hi) {
    mpz_t acc, c;
    mpz_init(acc); mpz_init(c);
    mpz_set_ui(acc, 1);
    for (size_t i = lo; i < hi; ++i) {
        imp_le(c, j->a + i * j->wa, j->wa);
        mpz_mul(acc, acc, c);
        mpz_mod(acc, acc, j->k1);
    }
    exp_le((uint8_t*)j->b, j->wout, acc);    /* b = this thread's slot in the partial-product array */
    mpz_clear(acc); mpz_clear(c);
}

int ref_add_reduce(const uint8_t* mod_be, size_t mod_len, size_t count, const uint8_t* c, size_t width,
                   uint8_t* out, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > count) threads = count ? (int)count : 1;
    mpz_t mod, acc, t;
    mpz_init(mod); mpz_init(acc); mpz_init(t);
    imp_be(mod, mod_be, mod_len);
    uint8_t* partial = (uint8_t*)calloc((size_t)threads, width);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
    for (int k = 0; k < threads; ++k) {
        jobs[k].fn = fold_items; jobs[k].a = c; jobs[k].wa = width; jobs[k].wout = width;
        jobs[k].b = partial + (size_t)k * width; jobs[k].out = partial;
        jobs[k].k1[0] = mod[0];
        jobs[k].lo = count * k / threads; jobs[k].hi = count * (k + 1) / threads;
        pthread_create(&th[k], NULL, trampoline, &jobs[k]);
    }
    for (int k = 0; k < threads; ++k) pthread_join(th[k], NULL);
    mpz_set_ui(acc, 1);
    for (int k = 0; k < threads; ++k) {
        imp_le(t, partial + (size_t)k * width, width);
        mpz_mul(acc, acc, t);
        mpz_mod(acc, acc, mod);
    }
    exp_le(out, width, acc);
    free(partial); free(th); free(jobs);
    mpz_clear(mod); mpz_clear(acc); mpz_clear(t);
    return 0;
}

/* ======================================================================================================
 * Round-2 additions: the threshold ZKP, the DDLEQ verifier, the encrypted dot product and the safe-prime
 * candidate procedure as libgmp call sequences (parity samples of >= 4096 items and CPU baselines beside
 * every GPU number).  SHA-256 comes from the image's libcrypto (crypto/sha256 in the reference).
 * ====================================================================================================== */
extern unsigned char* SHA256(const unsigned char* d, size_t n, unsigned char* md);
extern int __gmpz_probab_prime_p(const mpz_t, int);
extern unsigned long __gmpz_fdiv_ui(const mpz_t, unsigned long);
extern size_t __gmpz_sizeinbase(const mpz_t, int);
extern int __gmpz_tstbit(const mpz_t, unsigned long);
#define mpz_probab_prime_p __gmpz_probab_prime_p
#define mpz_fdiv_ui __gmpz_fdiv_ui
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_tstbit __gmpz_tstbit

/* gmp.Int.Bytes(): minimal big-endian magnitude, zero -> empty; appended to buf, returns the new length */
static size_t append_bytes(uint8_t** buf, size_t len, size_t* cap, const mpz_t z) {
    size_t cnt = 0;
    uint8_t* tmp = (uint8_t*)__gmpz_export(NULL, &cnt, 1, 1, 1, 0, z);
    if (len + cnt > *cap) { *cap = 2 * (len + cnt) + 64; *buf = (uint8_t*)realloc(*buf, *cap); }
    if (tmp) { memcpy(*buf + len, tmp, cnt); free(tmp); }
    return len + cnt;
}

typedef struct {
    size_t lo, hi, count;
    const uint8_t *c, *r, *dec_in, *e_in, *z_in;
    uint8_t *dec, *e, *z, *ok;
    size_t w2, wz;
    mpz_t n, share, v, vi;
    unsigned long l;
} zkp_job;

/* PartialDecryptionWithZKP (thresholdkey.go:225-255) with r supplied; computeHash :319-326, computeZ :313-317 */
static void* zkp_prove_items(void* p) {
    zkp_job* j = (zkp_job*)p;
    mpz_t c, r, delta, ex, n2, dec, c4, a, b, ci2, E, Z, four, two;
    mpz_init(c); mpz_init(r); mpz_init(delta); mpz_init(ex); mpz_init(n2); mpz_init(dec); mpz_init(c4); mpz_init(a); mpz_init(b);
    mpz_init(ci2); mpz_init(E); mpz_init(Z); mpz_init(four); mpz_init(two);
    mpz_set_ui(four, 4); mpz_set_ui(two, 2);
    uint8_t* buf = NULL; size_t cap = 0; uint8_t md[32];
    for (size_t i = j->lo; i < j->hi; ++i) {
        imp_le(c, j->c + i * j->w2, j->w2);
        imp_le(r, j->r + i * j->w2, j->w2);
        mpz_set_ui(delta, 1);
        for (unsigned long k = 1; k <= j->l; ++k) mpz_mul_ui(delta, delta, k);
        mpz_mul_ui(ex, delta, 2); mpz_mul(ex, j->share, ex);
        mpz_mul(n2, j->n, j->n);
        gmp_exp(dec, c, ex, n2);                               /* pd.Decryption = PartialDecrypt(c)      :230 */
        mpz_pow_ui(c4, c, 4);                                  /* c4 := Exp(c, 4, nil), unreduced         :241 */
        gmp_exp(a, c4, r, n2);                                 /* a := Exp(c4, r, n2)                     :242 */
        gmp_exp(b, j->v, r, n2);                               /* b := Exp(VerificationKey, r, n2)        :245 */
        mpz_pow_ui(ci2, dec, 2);                               /* ci2 := Exp(Decryption, 2, nil)          :248 */
        size_t len = 0;
        len = append_bytes(&buf, len, &cap, a); len = append_bytes(&buf, len, &cap, b);
        len = append_bytes(&buf, len, &cap, c4); len = append_bytes(&buf, len, &cap, ci2);
        SHA256(buf, len, md);
        imp_be(E, md, 32);                                     /* E = SetBytes(hash)                      :325 */
        mpz_mul(Z, E, delta); mpz_mul(Z, Z, j->share); mpz_add(Z, r, Z);   /* Z = r + E*delta*share      :313-317 */
        exp_le(j->dec + i * j->w2, j->w2, dec);
        exp_le(j->e + i * 32, 32, E);
        exp_le(j->z + i * j->wz, j->wz, Z);
    }
    free(buf);
    mpz_clear(c); mpz_clear(r); mpz_clear(delta); mpz_clear(ex); mpz_clear(n2); mpz_clear(dec); mpz_clear(c4); mpz_clear(a); mpz_clear(b);
    mpz_clear(ci2); mpz_clear(E); mpz_clear(Z); mpz_clear(four); mpz_clear(two);
    return NULL;
}

/* VerifyProof (thresholdkey.go:278-311) for proofs of one server (verification key vi) */
static void* zkp_verify_items(void* p) {
    zkp_job* j = (zkp_job*)p;
    mpz_t c, dec, E, Z, n2, c4, d2, a1, a2, a, b1, b2, b, exp_e;
    mpz_init(c); mpz_init(dec); mpz_init(E); mpz_init(Z); mpz_init(n2); mpz_init(c4); mpz_init(d2); mpz_init(a1); mpz_init(a2); mpz_init(a);
    mpz_init(b1); mpz_init(b2); mpz_init(b); mpz_init(exp_e);
    uint8_t* buf = NULL; size_t cap = 0; uint8_t md[32];
    for (size_t i = j->lo; i < j->hi; ++i) {
        imp_le(c, j->c + i * j->w2, j->w2);
        imp_le(dec, j->dec_in + i * j->w2, j->w2);
        imp_le(E, j->e_in + i * 32, 32);
        imp_le(Z, j->z_in + i * j->wz, j->wz);
        mpz_mul(n2, j->n, j->n);
        mpz_pow_ui(c4, c, 4); mpz_pow_ui(d2, dec, 2);          /* verifyPart1 :293-302 */
        gmp_exp(a1, c4, Z, n2);
        gmp_exp(a2, d2, E, n2);
        int good = mpz_invert(a2, a2, n2) != 0;
        mpz_mul(a, a1, a2); mpz_mod(a, a, n2);
        gmp_exp(b1, j->v, Z, n2);                              /* verifyPart2 :304-311 */
        gmp_exp(b2, j->vi, E, n2);
        good &= mpz_invert(b2, b2, n2) != 0;
        mpz_mul(b, b1, b2); mpz_mod(b, b, n2);
        size_t len = 0;
        len = append_bytes(&buf, len, &cap, a); len = append_bytes(&buf, len, &cap, b);
        len = append_bytes(&buf, len, &cap, c4); len = append_bytes(&buf, len, &cap, d2);
        SHA256(buf, len, md);
        imp_be(exp_e, md, 32);
        j->ok[i] = (uint8_t)(good && mpz_cmp(E, exp_e) == 0);  /* :290 */
    }
    free(buf);
    mpz_clear(c); mpz_clear(dec); mpz_clear(E); mpz_clear(Z); mpz_clear(n2); mpz_clear(c4); mpz_clear(d2); mpz_clear(a1); mpz_clear(a2); mpz_clear(a);
    mpz_clear(b1); mpz_clear(b2); mpz_clear(b); mpz_clear(exp_e);
    return NULL;
}

static void zkp_run(zkp_job* proto, void* (*fn)(void*), size_t count, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > count) threads = count ? (int)count : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    zkp_job* jobs = (zkp_job*)malloc(sizeof(zkp_job) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = *proto;
        jobs[t].lo = count * t / threads; jobs[t].hi = count * (t + 1) / threads;
        pthread_create(&th[t], NULL, fn, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

int ref_pdec_zkp(const uint8_t* n_be, size_t n_len, const uint8_t* share_be, size_t share_len, int l, const uint8_t* v_be, size_t v_len,
                 size_t count, const uint8_t* c, const uint8_t* r, size_t w_n2, uint8_t* dec, uint8_t* e, uint8_t* z, size_t w_z, int threads) {
    zkp_job j; memset(&j, 0, sizeof j);
    mpz_init(j.n); mpz_init(j.share); mpz_init(j.v); mpz_init(j.vi);
    imp_be(j.n, n_be, n_len); imp_be(j.share, share_be, share_len); imp_be(j.v, v_be, v_len);
    j.l = (unsigned long)l; j.c = c; j.r = r; j.dec = dec; j.e = e; j.z = z; j.w2 = w_n2; j.wz = w_z;
    zkp_run(&j, zkp_prove_items, count, threads);
    mpz_clear(j.n); mpz_clear(j.share); mpz_clear(j.v); mpz_clear(j.vi);
    return 0;
}

int ref_zkp_verify(const uint8_t* n_be, size_t n_len, const uint8_t* v_be, size_t v_len, const uint8_t* vi_be, size_t vi_len,
                   size_t count, const uint8_t* c, const uint8_t* dec, const uint8_t* e, const uint8_t* z, size_t w_n2, size_t w_z,
                   uint8_t* ok, int threads) {
    zkp_job j; memset(&j, 0, sizeof j);
    mpz_init(j.n); mpz_init(j.share); mpz_init(j.v); mpz_init(j.vi);
    imp_be(j.n, n_be, n_len); imp_be(j.v, v_be, v_len); imp_be(j.vi, vi_be, vi_len);
    j.c = c; j.dec_in = dec; j.e_in = e; j.z_in = z; j.ok = ok; j.w2 = w_n2; j.wz = w_z;
    zkp_run(&j, zkp_verify_items, count, threads);
    mpz_clear(j.n); mpz_clear(j.share); mpz_clear(j.v); mpz_clear(j.vi);
    return 0;
}

/* ---- verifyDDLEQProofInstance (ddleq.go:129-153): instance i belongs to statement i / secpar ---- */
typedef struct {
    size_t lo, hi; unsigned secpar;
    const uint8_t *ct1, *ct2, *x, *y, *alpha, *e, *f; uint8_t* ok;
    size_t wn, w2, w3;
    mpz_t n;
} ddleq_job;

static void* ddleq_verify_items(void* p) {
    ddleq_job* j = (ddleq_job*)p;
    mpz_t n2, n3, c1, c2, x, y, al, E, F, en, fn2, check, bit;
    mpz_init(n2); mpz_init(n3); mpz_init(c1); mpz_init(c2); mpz_init(x); mpz_init(y); mpz_init(al); mpz_init(E); mpz_init(F);
    mpz_init(en); mpz_init(fn2); mpz_init(check); mpz_init(bit);
    mpz_mul(n2, j->n, j->n); mpz_mul(n3, n2, j->n);
    uint8_t* buf = NULL; size_t cap = 0; uint8_t md[32];
    for (size_t i = j->lo; i < j->hi; ++i) {
        const size_t s = i / j->secpar;
        imp_le(c1, j->ct1 + s * j->w3, j->w3); imp_le(c2, j->ct2 + s * j->w3, j->w3);
        imp_le(x, j->x + i * j->wn, j->wn); imp_le(y, j->y + i * j->wn, j->wn);
        imp_le(al, j->alpha + i * j->w3, j->w3); imp_le(E, j->e + i * j->w2, j->w2); imp_le(F, j->f + i * j->w3, j->w3);
        /* RandomOracleBit(ct1, ct2, x, y, alpha): the first argument is skipped (random_oracle.go:24-26) */
        size_t len = 0;
        len = append_bytes(&buf, len, &cap, c2); len = append_bytes(&buf, len, &cap, x);
        len = append_bytes(&buf, len, &cap, y); len = append_bytes(&buf, len, &cap, al);
        SHA256(buf, len, md);
        const int chal = md[31] & 1;                            /* SetBytes(digest) mod 2 */
        mpz_set(check, chal ? c2 : c1);                         /* :140-143 */
        gmp_exp(en, E, j->n, n2);                               /* en := Exp(E, n, n2)   :145 */
        gmp_exp(fn2, F, n2, n3);                                /* fn2 := Exp(F, n2, n3) :146 */
        gmp_exp(check, check, en, n3);                          /* :148-150 */
        mpz_mul(check, check, fn2); mpz_mod(check, check, n3);
        j->ok[i] = (uint8_t)(mpz_cmp(al, check) == 0);
    }
    free(buf);
    mpz_clear(n2); mpz_clear(n3); mpz_clear(c1); mpz_clear(c2); mpz_clear(x); mpz_clear(y); mpz_clear(al); mpz_clear(E); mpz_clear(F);
    mpz_clear(en); mpz_clear(fn2); mpz_clear(check); mpz_clear(bit);
    return NULL;
}

int ref_ddleq_verify(const uint8_t* n_be, size_t n_len, size_t count, unsigned secpar, const uint8_t* ct1, const uint8_t* ct2,
                     const uint8_t* x, const uint8_t* y, const uint8_t* alpha, const uint8_t* e, const uint8_t* f,
                     size_t w_n, size_t w_n2, size_t w_n3, uint8_t* ok, int threads) {
    const size_t total = count * secpar;
    if (threads < 1) threads = 1;
    if ((size_t)threads > total) threads = total ? (int)total : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    ddleq_job* jobs = (ddleq_job*)calloc((size_t)threads, sizeof(ddleq_job));
    for (int t = 0; t < threads; ++t) {
        ddleq_job* j = &jobs[t];
        mpz_init(j->n); imp_be(j->n, n_be, n_len);
        j->secpar = secpar; j->ct1 = ct1; j->ct2 = ct2; j->x = x; j->y = y; j->alpha = alpha; j->e = e; j->f = f; j->ok = ok;
        j->wn = w_n; j->w2 = w_n2; j->w3 = w_n3;
        j->lo = total * t / threads; j->hi = total * (t + 1) / threads;
        pthread_create(&th[t], NULL, ddleq_verify_items, j);
    }
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); mpz_clear(jobs[t].n); }
    free(th); free(jobs);
    return 0;
}

/* ---- encrypted dot product: ConstMult (operations.go:58-64) per term, then Add (operations.go:11-29) ---- */
static void dot_items(const job_t* j, size_t lo, size_t hi) {
    mpz_t acc, c, k, t;
    mpz_init(acc); mpz_init(c); mpz_init(k); mpz_init(t);
    mpz_set_ui(acc, 1);
    for (size_t i = lo; i < hi; ++i) {
        imp_le(c, j->a + i * j->wa, j->wa);
        imp_le(k, j->out + i * 8, 8);                           /* out = the scalar array (read-only here) */
        gmp_exp(t, c, k, j->k1);                                /* ConstMult: Exp(ct.C, k, n2) */
        mpz_mul(acc, acc, t); mpz_mod(acc, acc, j->k1);         /* Add: accumulator = Mod(Mul(accumulator, c), n2) */
    }
    exp_le((uint8_t*)j->b, j->wout, acc);
    mpz_clear(acc); mpz_clear(c); mpz_clear(k); mpz_clear(t);
}

int ref_dot_u64(const uint8_t* mod_be, size_t mod_len, size_t count, const uint8_t* c, size_t width, const uint64_t* k,
                uint8_t* out, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > count) threads = count ? (int)count : 1;
    mpz_t mod, acc, t;
    mpz_init(mod); mpz_init(acc); mpz_init(t);
    imp_be(mod, mod_be, mod_len);
    uint8_t* partial = (uint8_t*)calloc((size_t)threads, width);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
    for (int q = 0; q < threads; ++q) {
        jobs[q].fn = dot_items; jobs[q].a = c; jobs[q].wa = width; jobs[q].wout = width;
        jobs[q].b = partial + (size_t)q * width; jobs[q].out = (uint8_t*)k;
        jobs[q].k1[0] = mod[0];
        jobs[q].lo = count * q / threads; jobs[q].hi = count * (q + 1) / threads;
        pthread_create(&th[q], NULL, trampoline, &jobs[q]);
    }
    for (int q = 0; q < threads; ++q) pthread_join(th[q], NULL);
    mpz_set_ui(acc, 1);
    for (int q = 0; q < threads; ++q) {
        imp_le(t, partial + (size_t)q * width, width);
        mpz_mul(acc, acc, t); mpz_mod(acc, acc, mod);
    }
    exp_le(out, width, acc);
    free(partial); free(th); free(jobs);
    mpz_clear(mod); mpz_clear(acc); mpz_clear(t);
    return 0;
}

/* ---- one iteration of runGenPrimeRoutine's loop per byte string (safe_prime.go:170-263); ProbablyPrime(20) is
 *      answered by mpz_probab_prime_p(q, 20) (Baillie-PSW + Miller-Rabin, like Go's).  CPU baseline of config 5. ---- */
static const unsigned SMALL_PRIMES[15] = {3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53};
static const unsigned long long SMALL_PRIMES_PRODUCT = 16294579238595022365ull;

static int prime_candidate(const mpz_t x, const mpz_t prod) {       /* isPrimeCandidate :280-290 */
    mpz_t m; mpz_init(m);
    mpz_mod(m, x, prod);
    unsigned long long mm = 0; size_t cnt = 0;
    __gmpz_export(&mm, &cnt, -1, 8, 0, 0, m);
    mpz_clear(m);
    for (int k = 0; k < 15; ++k) if (mm % SMALL_PRIMES[k] == 0 && mm != SMALL_PRIMES[k]) return 0;
    return 1;
}

typedef struct { size_t lo, hi; unsigned p_bits; const uint8_t* raw; uint8_t* ok; } sp_job;

static void* sp_items(void* p_) {
    sp_job* j = (sp_job*)p_;
    const unsigned qbits = j->p_bits - 1;
    unsigned b = qbits % 8; if (b == 0) b = 8;
    const size_t nb = (qbits + 7) / 8;
    uint8_t* bytes = (uint8_t*)malloc(nb);
    mpz_t p, q, prod, bm, two, pm1, t;
    mpz_init(p); mpz_init(q); mpz_init(prod); mpz_init(bm); mpz_init(two); mpz_init(pm1); mpz_init(t);
    __gmpz_import(prod, 1, -1, 8, 0, 0, &SMALL_PRIMES_PRODUCT);
    mpz_set_ui(two, 2);
    for (size_t i = j->lo; i < j->hi; ++i) {
        memcpy(bytes, j->raw + i * nb, nb);
        bytes[0] &= (uint8_t)((1u << b) - 1);
        if (b >= 2) bytes[0] |= (uint8_t)(3u << (b - 2));
        else { bytes[0] |= 1; if (nb > 1) bytes[1] |= 0x80; }
        bytes[nb - 1] |= 1;
        imp_be(q, bytes, nb);
        mpz_mod(bm, q, prod);
        unsigned long long mod = 0; size_t cnt = 0;
        __gmpz_export(&mod, &cnt, -1, 8, 0, 0, bm);
        for (unsigned long long delta = 0; delta < (1ull << 20); delta += 2) {
            const unsigned long long m = mod + delta;
            int sieved = 0;
            for (int k = 0; k < 15 && !sieved; ++k) if (m % SMALL_PRIMES[k] == 0 && (qbits > 6 || m != SMALL_PRIMES[k])) sieved = 1;
            if (sieved) continue;
            if (delta > 0) mpz_add_ui(q, q, (unsigned long)delta);       /* cumulative q += delta, as in the reference :216-219 */
            if (mpz_fdiv_ui(q, 3) == 1) continue;
            mpz_mul_ui(p, q, 2); mpz_add_ui(p, p, 1);
            if (!prime_candidate(p, prod)) continue;
            break;
        }
        int ok = mpz_probab_prime_p(q, 20) != 0;
        if (ok) { mpz_sub_ui(pm1, p, 1); mpz_powm(t, two, pm1, p); ok = mpz_cmp_ui(t, 1) == 0; }
        if (ok) ok = mpz_sizeinbase(q, 2) == qbits;
        j->ok[i] = (uint8_t)ok;
    }
    free(bytes);
    mpz_clear(p); mpz_clear(q); mpz_clear(prod); mpz_clear(bm); mpz_clear(two); mpz_clear(pm1); mpz_clear(t);
    return NULL;
}

int ref_safe_prime_scan(unsigned p_bits, size_t count, const uint8_t* raw, uint8_t* ok, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > count) threads = count ? (int)count : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    sp_job* jobs = (sp_job*)calloc((size_t)threads, sizeof(sp_job));
    for (int t = 0; t < threads; ++t) {
        jobs[t].p_bits = p_bits; jobs[t].raw = raw; jobs[t].ok = ok;
        jobs[t].lo = count * t / threads; jobs[t].hi = count * (t + 1) / threads;
        pthread_create(&th[t], NULL, sp_items, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
