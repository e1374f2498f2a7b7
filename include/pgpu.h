/*
 * pgpu.h -- C ABI of libpaillier_b200.so, the B200 batch engine behind the
 * hot path of sachaservan/paillier (modular exponentiation mod n^2, n^3, p^2, q^2).
 *
 * The reference has no FFI of its own for this path: its boundary is the
 * gmp.Int method set of github.com/ncw/gmp (one cgo call per big-integer
 * operation).  Each entry point below replaces N such call sequences by one
 * batched call; the reference call sites are cited per function.
 *
 * Data layout: every big integer crosses the boundary as a fixed-width record
 * of little-endian 32-bit limbs (little-endian bytes on the host).  Record
 * widths are per key and reported by pgpu_ctx_widths():
 *   n-width   : plaintexts m, randomness r, CRT outputs          (W_N  bytes)
 *   n2-width  : level-1 ciphertexts, partial decryptions, ZKP a,b (W_N2 = 2*W_N bytes)
 *   n3-width  : level-2 ciphertexts                               (W_N3 bytes, 0 if unsupported)
 * Record i of a batch starts at byte i*width.  Key material is passed as
 * big-endian magnitude bytes exactly as gmp.Int.Bytes() returns it.
 *
 * Ownership: the caller owns every buffer; the library keeps no host pointer
 * after a call returns.  Calls block until the result is in the output buffer.
 * A context is bound to one device and must be used by one thread at a time.
 * All functions return PGPU_OK (0) or an error code; pgpu_last_error() gives
 * the message.  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with PGPU_ERR_CUDA.
 */
#ifndef PGPU_H
#define PGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgpu_ctx pgpu_ctx;

enum {
    PGPU_OK = 0,
    PGPU_ERR_ARG = 1,            /* bad argument (null pointer, width mismatch, even modulus ...) */
    PGPU_ERR_CUDA = 2,           /* CUDA runtime error or no device */
    PGPU_ERR_NCCL = 3,
    PGPU_ERR_STATE = 4,          /* key material for this call was not loaded into the context */
    PGPU_ERR_NOT_INVERTIBLE = 5, /* ModInverse of a non-unit (undefined in the reference) */
    PGPU_ERR_THRESHOLD = 6,      /* "Threshold not meet" / duplicate share ids, thresholdkey.go:77-89 */
    PGPU_ERR_UNSUPPORTED = 7     /* key size outside the built kernel shapes */
};

/* modulus selectors for the generic entry points */
enum { PGPU_MOD_N = 0, PGPU_MOD_N2 = 1, PGPU_MOD_N3 = 2 };

int pgpu_version(void);
int pgpu_device_count(int* count);
/* message of the last failing call on this thread (ctx may be NULL) */
const char* pgpu_last_error(const pgpu_ctx* ctx);

/* ---- key state -------------------------------------------------------- */

/* PublicKey{N} (paillier.go:46-56); g = n+1 is implied (paillier.go:147).
 * Precomputes n^2, n^3 (GetN2/GetN3, paillier.go:72-90) and Montgomery constants. */
int pgpu_ctx_create(pgpu_ctx** out, int device, const uint8_t* n_be, size_t n_len);
int pgpu_ctx_destroy(pgpu_ctx* ctx);
/* record widths in bytes (any pointer may be NULL) */
int pgpu_ctx_widths(const pgpu_ctx* ctx, size_t* w_n, size_t* w_n2, size_t* w_n3);
/* record width in bytes of the generic entry points (pgpu_modexp, pgpu_modmul, pgpu_modinv ...) for one modulus selector;
 * equals the n2- / n3-width above for PGPU_MOD_N2 / PGPU_MOD_N3, 0 if that modulus is not available */
int pgpu_ctx_mod_width(const pgpu_ctx* ctx, int modsel, size_t* width);
/* run subsequent *_dev calls on this cudaStream_t (default: a stream owned by the context) */
int pgpu_ctx_set_stream(pgpu_ctx* ctx, void* cuda_stream);

/* SecretKey (paillier.go:59-62).  The reference keeps only Lambda = (p-1)(q-1)
 * (paillier.go:152,172-176); pgpu_ctx_set_secret_lambda recovers p, q from
 * (n, lambda).  Either call enables CRT decryption over p^2, q^2. */
int pgpu_ctx_set_secret_pq(pgpu_ctx* ctx, const uint8_t* p_be, size_t p_len, const uint8_t* q_be, size_t q_len);
int pgpu_ctx_set_secret_lambda(pgpu_ctx* ctx, const uint8_t* lambda_be, size_t lambda_len);

/* ThresholdSecretKey / ThresholdPublicKey (thresholdkey.go:26-46).
 * share may be NULL for a verifier/combiner-only context.  vkeys = l
 * verification keys v_i as n2-width records (may be NULL). */
int pgpu_ctx_set_threshold(pgpu_ctx* ctx, int total_servers, int threshold, int id,
                           const uint8_t* share_be, size_t share_len,
                           const uint8_t* v_be, size_t v_len,
                           const void* vkeys_n2w);

/* ---- batch entry points, host buffers ---------------------------------- */

/* PublicKey.EncryptWithR (paillier.go:185-187,206-218), level 1:
 * c[i] = (1 + m[i]*n) * r[i]^n mod n^2.   m, r: n-width; c: n2-width. */
int pgpu_encrypt_with_r(pgpu_ctx* ctx, size_t count, const void* m, const void* r, void* c);

/* EncryptWithR called on a SecretKey (the embedded PublicKey's method, paillier.go:29-34,185-187,206-218):
 * the same c as pgpu_encrypt_with_r, bit for bit, but r^n is computed mod p^2 and mod q^2 and recombined
 * (two half-width exponentiations, about 2x the throughput).  Needs pgpu_ctx_set_secret. */
int pgpu_encrypt_with_r_sk(pgpu_ctx* ctx, size_t count, const void* m, const void* r, void* c);

/* Offline/online EncryptWithR (the r^n pool): rn[i] = r[i]^n mod n^2 is prepared ahead of time --
 * pgpu_encrypt_with_r (or _sk) with m = 0 returns exactly that -- and the online call is two multiplications:
 * c[i] = (1 + m[i]*n) * rn[i] mod n^2, the same c as EncryptWithR(m[i], r[i]) (paillier.go:206-218).
 * m: n-width; rn, c: n2-width.  Each rn must be used once (it is the ciphertext's randomness). */
int pgpu_encrypt_with_rn(pgpu_ctx* ctx, size_t count, const void* m, const void* rn, void* c);

/* SecretKey.Decrypt (paillier.go:292-303), level 1, computed with CRT over
 * p^2 and q^2.   c: n2-width; m: n-width. */
int pgpu_decrypt(pgpu_ctx* ctx, size_t count, const void* c, void* m);

/* PublicKey.ConstMult (operations.go:58-64): out[i] = c[i]^k[i] mod n^2 with
 * k an unsigned record of k_bytes bytes (multiple of 4).  k = 0 yields 1. */
int pgpu_const_mult(pgpu_ctx* ctx, size_t count, const void* c, const void* k, size_t k_bytes, void* out);

/* PublicKey.Add over a whole batch (operations.go:11-29): out = prod c[i] mod n^2. */
int pgpu_add_reduce(pgpu_ctx* ctx, size_t count, const void* c, void* out);
/* Add over a batch of ciphertexts of `level` (1: mod n^2, 2: mod n^3; operations.go:11-29 takes the modulus of
 * cts[0].Level): records of the n2- / n3-width */
int pgpu_add_reduce_at_level(pgpu_ctx* ctx, int level, size_t count, const void* c, void* out);
/* PublicKey.Add(a[i], b[i]) element-wise: out[i] = a[i]*b[i] mod n^2. */
int pgpu_add_pairs(pgpu_ctx* ctx, size_t count, const void* a, const void* b, void* out);
/* Encrypted dot product: out = prod c[i]^k[i] mod n^2 (ConstMult + Add fused). */
int pgpu_dot_u64(pgpu_ctx* ctx, size_t count, const void* c, const uint64_t* k, void* out);

/* ThresholdSecretKey.PartialDecrypt (thresholdkey.go:192-201):
 * out[i] = c[i]^(2*delta*share) mod n^2, delta = l!.   c, out: n2-width. */
int pgpu_partial_decrypt(pgpu_ctx* ctx, size_t count, const void* c, void* out);

/* PublicKey.Sub(a[i], b[i]) (operations.go:32-55): out[i] = a[i] * b[i]^-1 mod n^2.
 * PGPU_ERR_NOT_INVERTIBLE names the first b[i] that is not a unit. */
int pgpu_sub_pairs(pgpu_ctx* ctx, size_t count, const void* a, const void* b, void* out);
/* gmp.Int.ModInverse (mpz_invert) against one of the key's moduli */
int pgpu_modinv(pgpu_ctx* ctx, int modsel, size_t count, const void* a, void* out);

/* ThresholdSecretKey.PartialDecryptionWithZKP (thresholdkey.go:225-255) with the random r in [0, n^2)
 * supplied by the caller (the reference draws it from crypto/rand at :233).
 * c, r, dec: n2-width; e: 32-byte records (SetBytes(sha256 digest), little-endian); z: z-width records
 * (pgpu_ctx_z_width), Z = r + E*delta*share unreduced (:313-317). */
int pgpu_pdec_zkp_prove(pgpu_ctx* ctx, size_t count, const void* c, const void* r, void* dec, void* e, void* z);
int pgpu_ctx_z_width(const pgpu_ctx* ctx, size_t* w_z);
/* PartialDecryptionZKP.VerifyProof (thresholdkey.go:278-311) for proofs of server `id`
 * (verification key VerificationKeys[id-1], :305); ok[i] = 1 if the proof verifies. */
int pgpu_pdec_zkp_verify(pgpu_ctx* ctx, size_t count, int id, const void* c, const void* dec, const void* e, const void* z, uint8_t* ok);
/* ThresholdPublicKey.CombinePartialDecryptions (thresholdkey.go:149-161) for a batch of ciphertexts:
 * k shares with server ids[0..k); decs holds k consecutive batches of `count` n2-width partial
 * decryptions (share j's batch starts at record j*count).  m: n-width plaintexts.
 * PGPU_ERR_THRESHOLD for "Threshold not meet" / duplicate ids (:77-89). */
int pgpu_combine(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, void* m);
/* host-buffer form of pgpu_combine_verified_dev (see there): decs = k batches of count records, ok = k * count verdict bytes */
int pgpu_combine_verified(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, const uint8_t* ok, void* m, uint8_t* item_ok);


/* ---- level 2 (mod n^3), alternative encryption, randomness, nested ops ---- */
/* `level` is 1 (EncLevelOne) or 2 (EncLevelTwo), paillier.go:17-23.  At level 2 plaintexts are
 * n2-width, ciphertexts n3-width; PGPU_ERR_UNSUPPORTED if n^3 exceeds the built kernel shapes. */

/* PublicKey.EncryptWithRAtLevel (paillier.go:206-218): c = (1+n)^m * r^(n^s) mod n^(s+1); r: n-width */
int pgpu_encrypt_with_r_at_level(pgpu_ctx* ctx, int level, size_t count, const void* m, const void* r, void* c);
/* the same ciphertexts for the holder of p, q (pgpu_ctx_set_secret): level 1 = pgpu_encrypt_with_r_sk; level 2 computes
 * r^(n^2) over p^3 and q^3 with the exponent reduced mod phi (+ phi, so that non-units stay exact) */
int pgpu_encrypt_with_r_at_level_sk(pgpu_ctx* ctx, int level, size_t count, const void* m, const void* r, void* c);
/* SecretKey.Decrypt at either level (paillier.go:292-340); level 2 runs recoveryAlgorithm(s = 2) */
int pgpu_decrypt_at_level(pgpu_ctx* ctx, int level, size_t count, const void* c, void* m);
/* PublicKey.H and K = 2^k_bits (paillier.go:46-56,151,158-161): enables AltEncrypt; precomputes the
 * generators h_1, h_2 (getGeneratorOfQuadraticResiduesForLevel, :416-434) and their fixed-base tables */
int pgpu_ctx_set_alt_generator(pgpu_ctx* ctx, const uint8_t* h_be, size_t h_len, unsigned k_bits);
/* PublicKey.AltEncryptWithRAtLevel (paillier.go:221-238): c = (1+n)^m * h_s^(r mod K) mod n^(s+1).
 * r: n-width; the reference reduces the caller's r mod K in place (:228) -- here r is read-only and
 * the binding applies the same reduction to the caller's value. */
int pgpu_alt_encrypt_with_r_at_level(pgpu_ctx* ctx, int level, size_t count, const void* m, const void* r, void* c);
/* PublicKey.Randomize (operations.go:67-69) with the r of the fresh Encrypt(0) supplied: out = c * r^n mod n^2 */
int pgpu_randomize_with_r(pgpu_ctx* ctx, size_t count, const void* c, const void* r, void* out);
/* SecretKey.ExtractRandonness (operations.go:75-91); out: n-width */
int pgpu_extract_randomness(pgpu_ctx* ctx, int level, size_t count, const void* c, void* out);
/* PublicKey.NestedRandomize (operations.go:96-118) with a, b (n-width) supplied: ct^(a^n mod n^2) * b^(n^2) mod n^3 */
int pgpu_nested_randomize_with(pgpu_ctx* ctx, size_t count, const void* ct, const void* a, const void* b, void* out);
/* PublicKey.NestedAdd / NestedSub (operations.go:121-140): ct1 level 2 (n3-width), ct2 level 1 (n2-width) */
int pgpu_nested_add(pgpu_ctx* ctx, size_t count, const void* ct1, const void* ct2, void* out);
int pgpu_nested_sub(pgpu_ctx* ctx, size_t count, const void* ct1, const void* ct2, void* out);

/* ---- DDLEQ proofs (ddleq.go) ---------------------------------------------- */
/* SecretKey.ProveDDLEQ (ddleq.go:27-127) for `count` statements (ct1, ct2, a, b) with `secpar` instances
 * each; instance j of statement i is record i*secpar + j.  x, y (n-width, in Z*_n) are the per-instance
 * randomness the reference draws at :71-79.  Outputs: alpha, f n3-width; e n2-width (X = x, Y = y).
 * PGPU_ERR_ARG "cannot prove re-encryption because inputs are wrong" where the reference panics (:67-69). */
int pgpu_ddleq_prove(pgpu_ctx* ctx, size_t count, unsigned secpar, const void* ct1, const void* ct2, const void* a, const void* b,
                     const void* x, const void* y, void* alpha, void* e, void* f);
/* PublicKey.VerifyDDLEQProof (ddleq.go:44-53,129-153): ok[i*secpar + j] = 1 if instance j of proof i
 * verifies; a proof is valid when all its instances are. */
int pgpu_ddleq_verify(pgpu_ctx* ctx, size_t count, unsigned secpar, const void* ct1, const void* ct2, const void* x, const void* y,
                      const void* alpha, const void* e, const void* f, uint8_t* ok);

/* ---- safe-prime candidate testing (safe_prime.go) ---------------------------- */
/* One iteration of runGenPrimeRoutine's loop (safe_prime.go:170-263) for each of `count` byte strings
 * as the reference reads them from its io.Reader: raw = records of ceil((p_bits-1)/8) bytes.  The
 * masks (:175-202), the sieve mod 3*5*...*53 with the delta search (:208-249, including the cumulative
 * q += delta), q.ProbablyPrime(20) (Miller-Rabin, bases 2..71), the base-2 Fermat test on p = 2q+1
 * (:272-278) and the bit-length check (:256-258) run on `device`.  p, q: records of
 * 4*S bytes, S = 32 / 48 / 64 limbs for p_bits <= 1024 / 1536 / 2048; ok[i] = 1 if candidate i is accepted.
 * These two calls need no key context; errors are reported by pgpu_primes_last_error(). */
int pgpu_safe_prime_scan(int device, unsigned p_bits, size_t count, const uint8_t* raw, void* p_out, void* q_out, uint8_t* ok,
                         uint64_t* launches);
/* big.Int.ProbablyPrime stand-in: `rounds` (1..20) Miller-Rabin strong tests, bases 2, 3, 5, ... on odd
 * candidates that all have exactly `bits` bits (same record layout as above). */
int pgpu_miller_rabin(int device, unsigned bits, size_t count, const void* cand, unsigned rounds, uint8_t* ok, uint64_t* launches);
const char* pgpu_primes_last_error(void);

/* Generic batched gmp.Int.Exp (mpz_powm) / Mul+Mod against one of the key's
 * moduli; records are the modulus' width.  exp: per-item unsigned records of
 * exp_bytes bytes.  These back ConstMult, the ZKP and DDLEQ entry points. */
int pgpu_modexp(pgpu_ctx* ctx, int modsel, size_t count, const void* base, const void* exp, size_t exp_bytes, void* out);
int pgpu_modexp_shared(pgpu_ctx* ctx, int modsel, size_t count, const void* base, const uint8_t* exp_be, size_t exp_len, void* out);
int pgpu_modmul(pgpu_ctx* ctx, int modsel, size_t count, const void* a, const void* b, void* out);

/* ---- device buffers and pinned host memory ------------------------------ */
/* SURVEY.md 8(b) "Ownership" / 8(f) rank 1: ciphertexts stay on the device between calls (Encrypt -> ConstMult -> Add ->
 * Decrypt without PCIe round trips, the callers of operations.go:11-64).  A pgpu_buf is device memory on the context's
 * device; pgpu_buf_ptr() is what the *_dev entry points below take.  Upload / download block until the copy is done (no
 * host pointer is retained); pgpu_buf_free waits for the work enqueued on the context before releasing the memory.  Free
 * every buffer before pgpu_ctx_destroy of the context it was allocated on. */
typedef struct pgpu_buf pgpu_buf;
int pgpu_buf_alloc(pgpu_ctx* ctx, size_t bytes, pgpu_buf** out);
int pgpu_buf_free(pgpu_buf* buf);
void* pgpu_buf_ptr(const pgpu_buf* buf);
size_t pgpu_buf_size(const pgpu_buf* buf);
int pgpu_buf_upload(pgpu_buf* dst, size_t dst_off, const void* host, size_t bytes);
int pgpu_buf_download(const pgpu_buf* src, size_t src_off, void* host, size_t bytes);
/* wait for everything enqueued by *_dev calls on this context */
int pgpu_ctx_sync(pgpu_ctx* ctx);
/* page-locked host memory: with it the host-buffer entry points copy at full PCIe rate and overlap the copies of one chunk
 * with the kernels of the next (Go: wrap the pointer with unsafe.Slice; pageable slices work too, at the pageable rate) */
int pgpu_host_alloc(size_t bytes, void** out);
int pgpu_host_free(void* p);

/* ---- same operations on device-resident buffers ------------------------ */
/* Pointers are device pointers on the context's device; work is enqueued on
 * the context's stream and NOT synchronised (the caller owns ordering). */
int pgpu_encrypt_with_r_dev(pgpu_ctx* ctx, size_t count, const void* m, const void* r, void* c);
int pgpu_encrypt_with_r_sk_dev(pgpu_ctx* ctx, size_t count, const void* m, const void* r, void* c);
int pgpu_encrypt_with_rn_dev(pgpu_ctx* ctx, size_t count, const void* m, const void* rn, void* c);
int pgpu_decrypt_dev(pgpu_ctx* ctx, size_t count, const void* c, void* m);
int pgpu_partial_decrypt_dev(pgpu_ctx* ctx, size_t count, const void* c, void* out);
int pgpu_const_mult_dev(pgpu_ctx* ctx, size_t count, const void* c, const void* k, size_t k_bytes, void* out);
int pgpu_add_reduce_dev(pgpu_ctx* ctx, size_t count, const void* c, void* out);
int pgpu_dot_u64_dev(pgpu_ctx* ctx, size_t count, const void* c, const uint64_t* k, void* out);
int pgpu_add_pairs_dev(pgpu_ctx* ctx, size_t count, const void* a, const void* b, void* out);
/* first_bad: device word set to 0xffffffff, or to the index of the first b[i] that is not a unit */
int pgpu_sub_pairs_dev(pgpu_ctx* ctx, size_t count, const void* a, const void* b, void* out, uint32_t* first_bad);
int pgpu_randomize_with_r_dev(pgpu_ctx* ctx, size_t count, const void* c, const void* r, void* out);

int pgpu_pdec_zkp_prove_dev(pgpu_ctx* ctx, size_t count, const void* c, const void* r, void* dec, void* e, void* z);
/* the proof half of PartialDecryptionWithZKP (thresholdkey.go:233-254) for partial decryptions the caller already holds
 * (dec = output of pgpu_partial_decrypt_dev for the same c): computes E and Z only, so that a share-holder can time or
 * pipeline PartialDecrypt and the proof separately.  Same E, Z as pgpu_pdec_zkp_prove_dev. */
int pgpu_pdec_zkp_prove_given_dev(pgpu_ctx* ctx, size_t count, const void* c, const void* r, const void* dec, void* e, void* z);
int pgpu_combine_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, void* m);
/* as pgpu_combine_dev, but share j's batch starts at record j*share_stride of decs: combines a slice of the
 * all-gathered [share][ciphertext] buffer in place (one share-holder per GPU) */
int pgpu_combine_strided_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, size_t share_stride, void* m);
int pgpu_pdec_zkp_verify_dev(pgpu_ctx* ctx, size_t count, int id, const void* c, const void* dec, const void* e, const void* z, uint8_t* ok);
/* CombinePartialDecryptionsZKP (thresholdkey.go:164-172) with the reference's PER-CIPHERTEXT filter: ok holds k * count
 * verdict bytes grouped by server like decs (ok[j*count + i] != 0: server ids[j]'s proof for ciphertext i verified); a
 * ciphertext is combined from the shares whose proof holds, one combine per distinct surviving set.  Ciphertexts left with
 * fewer than `threshold` valid shares get a zero plaintext and item_ok[i] = 0 (item_ok may be NULL); if there are any the
 * call returns PGPU_ERR_THRESHOLD ("Threshold not meet") AFTER filling every other plaintext. */
int pgpu_combine_verified_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, size_t share_stride, const uint8_t* ok,
                              void* m, uint8_t* item_ok);
/* VerifyProof for the proofs of k servers in one batch: n_per_id proofs per server, grouped by server in the order of
 * ids[0..k) (record i belongs to server ids[i / n_per_id]); c, dec, e, z, ok hold k * n_per_id records.  Used by the
 * combiner of a threshold round, which checks every share-holder's proofs of its ciphertext slice at once. */
int pgpu_pdec_zkp_verify_multi_dev(pgpu_ctx* ctx, size_t n_per_id, int k, const int* ids, const void* c, const void* dec, const void* e,
                                   const void* z, uint8_t* ok);

/* ---- several devices of one process: BASELINE config 4 (SURVEY.md 8b "Threading", 8e) -------------- */
/* One context per device, each holding one share of the SAME threshold key (pgpu_ctx_set_threshold with its id and
 * share).  pgpu_multi_create builds one NCCL communicator per device (ncclCommInitAll; NCCL is loaded at this call with
 * dlopen("libnccl.so.2") and reported as PGPU_ERR_NCCL when absent).  The handle borrows the contexts: destroy it first. */
typedef struct pgpu_multi pgpu_multi;
int pgpu_multi_create(pgpu_multi** out, pgpu_ctx* const* ctxs, int n);
int pgpu_multi_destroy(pgpu_multi* m);
int pgpu_multi_size(const pgpu_multi* m);
const char* pgpu_multi_last_error(const pgpu_multi* m);
/* One threshold-decryption round over host buffers: c = count n2-width ciphertext records (read by every device);
 * zkp_r = NULL (PartialDecrypt only, thresholdkey.go:192-201 + :149-161) or one pointer per context to count n2-width r
 * records (PartialDecryptionWithZKP :225-255, VerifyProof :278-311, CombinePartialDecryptionsZKP :164-172 with its
 * per-ciphertext filter).  Device g computes its share's partial decryptions (and proofs) of ALL ciphertexts, one
 * ncclAllGather per field lands [share][ciphertext] on every device, device g verifies and combines ciphertext slice g.
 * plain = count n-width records; item_ok (may be NULL) = 0 where fewer than `threshold` proofs verified, in which case the
 * call returns PGPU_ERR_THRESHOLD after filling the rest.  Blocking; one host thread per device inside. */
int pgpu_multi_threshold_round(pgpu_multi* m, size_t count, const void* c, const void* const* zkp_r, void* plain, uint8_t* item_ok);
/* device milliseconds of the last round per phase (max over the devices): pdec, prove, all-gather, verify, combine */
int pgpu_multi_last_phases_ms(const pgpu_multi* m, float* out5);

/* VerifyProof of k <= 8 servers' proofs for the SAME n ciphertexts (what the combiner of a threshold round checks): c = n
 * records; dec, e, z, ok = n * k records ITEM-MAJOR (record i*k + j is server ids[j]'s proof for ciphertext i).  Same verdicts
 * as pgpu_pdec_zkp_verify_multi_dev; the k powers (c^4)^Z of one ciphertext share their squarings, 3.7x fewer multiplications
 * for them at k = 8. */
int pgpu_pdec_zkp_verify_shared_dev(pgpu_ctx* ctx, size_t n, int k, const int* ids, const void* c, const void* dec, const void* e, const void* z,
                                    uint8_t* ok);

/* ---- introspection used by bench.py ------------------------------------ */
/* number of kernels this context has launched so far */
int pgpu_ctx_launch_count(const pgpu_ctx* ctx, uint64_t* launches);
/* Montgomery multiplications per item of the compiled programs
 * (what = 0 encrypt, 1 decrypt (both CRT halves), 2 partial decrypt, 3 secret-key encrypt (both halves)): squarings and multiplies */
int pgpu_ctx_program_cost(const pgpu_ctx* ctx, int what, uint32_t* limbs, uint32_t* n_sqr, uint32_t* n_mul);
/* which exponentiation kernel serves a modulus (modsel: 0 n, 1 n^2, 2 n^3, 3 p^2 / q^2, 4 p^3 / q^3): lanes per residue,
 * limbs per lane, fp64 = 1 for the FP64-pipe kernel (52-bit limbs, powm_vm52) / 0 for the integer-pipe kernel (32-bit
 * limbs, powm_vm), and how many residues a full persistent grid holds (any pointer may be NULL) */
int pgpu_ctx_kernel_shape(const pgpu_ctx* ctx, int modsel, int* tpi, int* limbs_per_lane, int* fp64, int* resident_groups);
/* device time of the last call's kernels in milliseconds (host-buffer and *_dev
 * calls record CUDA events around their launches when timing is enabled) */
int pgpu_ctx_enable_timing(pgpu_ctx* ctx, int on);
int pgpu_ctx_last_kernel_ms(pgpu_ctx* ctx, float* ms);

/* host big-integer self test hook (tests/test_host_bignum.py): op 0 mul, 1 divmod-q,
 * 2 mod, 3 modinv, 4 modexp, 5 isqrt.  Operands and result are big-endian bytes;
 * *out_len holds the capacity on entry and the length on return. */
int pgpu_selftest_bn(int op, const uint8_t* a, size_t a_len, const uint8_t* b, size_t b_len,
                     const uint8_t* m, size_t m_len, uint8_t* out, size_t* out_len);

/* host interpreter of the exponentiation programs the library compiles, for the CPU test suite (no GPU, not used by any
 * other entry point): kind 0 = sliding window over the shared exponent shared_exp_be -> base^e; 1 = fixed windows over the
 * per-item exponent item_exps[0..exp_limbs) -> base^exp; 2 = shared-base multi-exponentiation of k exponents (item_exps =
 * k records of exp_limbs little-endian limbs) -> (base^pre)^exp_s; 3 = PartialDecrypt fused with the proof's power ->
 * base^e and (base^4)^exp.  out = *n_out big-endian records of mod_len bytes; n_sqr / n_mul = the program's cost. */
int pgpu_selftest_program(int kind, const uint8_t* mod_be, size_t mod_len, const uint8_t* base_be, size_t base_len,
                          const uint8_t* shared_exp_be, size_t shared_len, const uint32_t* item_exps, uint32_t exp_limbs, uint32_t k, uint32_t pre,
                          uint8_t* out, size_t out_cap, uint32_t* n_out, uint32_t* n_sqr, uint32_t* n_mul);

/* the micro-program (csrc/vm.h ops, OP_END included) pgpu_selftest_program compiled last on the calling thread and the number of
 * table entries it uses: tests/test_mont_host_emulation.py runs it through the interpreter's SOURCE (csrc/vm_run.cuh) on an
 * emulated warp of CPU threads.  *n_ops is set even when the buffer is too small.  Test hook only. */
int pgpu_selftest_last_program(uint32_t* ops, size_t cap, size_t* n_ops, uint32_t* tbl_entries);

#ifdef __cplusplus
}
#endif
#endif /* PGPU_H */
