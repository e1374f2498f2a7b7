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
