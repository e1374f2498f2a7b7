// Level-2 (mod n^3) encryption and decryption, alternative encryption with fixed-base tables,
// randomness extraction, nested operations and the DDLEQ proofs, assembled from the engine's
// device-side operations (engine.hpp).  Reference: paillier.go:206-238,292-372,
// operations.go:67-140, ddleq.go:27-153, random_oracle.go:10-32.
#include "engine.hpp"

namespace pgpu {

namespace {

BigU to_mont(const ModCtx& M, const BigU& x) { return ((x % M.N) * M.R1) % M.N; }

// x = m*R on entry (Montgomery form); afterwards x = (1+n)^m * R mod n^2 = (1 + m*n) * R
void emit_g_pow_level1(Program& P) {
    P.emit(OP_MULC, K_NM); P.n_mul++;
    P.emit(OP_ADDC, K_R1);
}

// x = m*R on entry; afterwards x = (1+n)^m * R mod n^3 = (1 + m*n + m(m-1)/2 * n^2) * R.
// Uses table slot `slot`.
void emit_g_pow_level2(Program& P, uint32_t slot) {
    P.emit(OP_STT, slot); P.use_slot(slot);
    P.emit(OP_ADDC, K_NEG1);                 // m - 1
    P.emit(OP_MULC, K_C1); P.n_mul++;        // (m - 1) * n^2/2
    P.emit(OP_ADDC, K_NM);                   // + n
    P.emit(OP_MULT, slot); P.n_mul++;        // * m
    P.emit(OP_ADDC, K_R1);                   // + 1
}

// x = product of base^(digit_k * 2^(w*k)) from the fixed table T (starts from 1)
void emit_pow_fixed(Program& P, const FixedTable& T) {
    P.emit(OP_LDC, K_R1);
    for (uint32_t k = 0; k < T.nwin; ++k) { P.emit(OP_FIXW, (k * T.w) | (T.w << 20)); P.n_mul++; }
}

}  // namespace

void fixed_table_free(FixedTable& T) { if (T.d) cudaFree(T.d); T = FixedTable(); }

// T = comb table of base^(d * 2^(w*k)) mod M for exponents of exp_bits bits, built on the device:
// rows base^(2^(w*k)) by one batched exponentiation, the 2^w multiples of each row by a second one.
int fixed_table_build(pgpu_ctx* ctx, const ModCtx& M, const BigU& base, uint32_t exp_bits, FixedTable& T) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    fixed_table_free(T);
    if (exp_bits == 0) return fail(ctx, PGPU_ERR_ARG, "fixed-base table for an empty exponent");
    const uint32_t S = M.sh.S;
    // widest window whose table stays below the cap (PGPU_FIXED_TABLE_MB, default 32 = L2-resident); one multiplication per window
    static const size_t cap_mb = [] { const char* e = getenv("PGPU_FIXED_TABLE_MB"); const long v = e ? atol(e) : 32; return (size_t)(v < 1 ? 1 : v); }();
    uint32_t w = 1;
    for (uint32_t c = 2; c <= 8; ++c) {
        const size_t nwin = (exp_bits + c - 1) / c;
        if ((nwin << c) * S * 4 <= (cap_mb << 20)) w = c;
    }
    const uint32_t nwin = (exp_bits + w - 1) / w, per = 1u << w;
    const uint32_t el = (w * (nwin - 1)) / 32 + 1;
    std::vector<uint32_t> e1((size_t)nwin * el, 0), e2((size_t)nwin * per);
    for (uint32_t k = 0; k < nwin; ++k) e1[(size_t)k * el + (w * k) / 32] = 1u << ((w * k) % 32);
    for (size_t i = 0; i < e2.size(); ++i) e2[i] = (uint32_t)(i % per);
    DEVBUF(de1, ctx, e1.size()); DEVBUF(de2, ctx, e2.size()); DEVBUF(db, ctx, S); DEVBUF(dr1, ctx, S);
    DEVBUF(rows, ctx, (size_t)nwin * S); DEVBUF(ent, ctx, (size_t)nwin * per * S);
    int rc;
    if ((rc = upload(ctx, de1.p, e1))) return rc;
    if ((rc = upload(ctx, de2.p, e2))) return rc;
    if ((rc = upload(ctx, db.p, (base % M.N).limbs(S)))) return rc;
    if ((rc = upload(ctx, dr1.p, M.R1.limbs(S)))) return rc;
    if ((rc = modexp_items_io(ctx, M, nwin, IoDesc{db.p, 0, S}, ExpDesc{de1.p, el, 32 * el, nullptr}, rows.p))) return rc;
    if ((rc = modexp_items_io(ctx, M, (size_t)nwin * per, IoDesc{rows.p, S, S, per}, ExpDesc{de2.p, 1, w, nullptr}, ent.p))) return rc;
    CU(ctx, cudaMalloc(&T.d, (size_t)nwin * per * S * 4));
    if ((rc = modmul_io(ctx, M, (size_t)nwin * per, IoDesc{ent.p, S, S}, IoDesc{dr1.p, 0, S}, T.d))) return rc;   // -> Montgomery form
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    T.w = w; T.nwin = nwin; T.bits = exp_bits;
    return PGPU_OK;
}

// out[i] = base^exp[i] mod M for the fixed base of table T (comb method: one multiplication per window)
int modexp_fixed_dev(pgpu_ctx* ctx, const ModCtx& M, const FixedTable& T, size_t count, const ExpDesc& exp, uint32_t* out) {
    if (!T.d) return fail(ctx, PGPU_ERR_STATE, "fixed-base table was not built");
    if (exp.bits > T.bits) return fail(ctx, PGPU_ERR_ARG, "exponent wider than the fixed-base table");
    const uint32_t nwin = (exp.bits + T.w - 1) / T.w;
    const std::string key = "powf:" + std::to_string(M.sh.S) + ":" + std::to_string(T.w) + ":" + std::to_string(nwin);
    Program* P = cached_program(ctx, key);
    if (!P) {
        Program np;
        np.emit(OP_LDC, K_R1);
        for (uint32_t k = 0; k < nwin; ++k) { np.emit(OP_FIXW, (k * T.w) | (T.w << 20)); np.n_mul++; }
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STO, 0);
        int rc = program_upload(ctx, np);
        if (rc) return rc;
        P = &(ctx->prog_cache[key] = np);
    }
    ExpDesc ex = exp; ex.fixed = T.d;
    return run_vm(ctx, M, *P, count, nullptr, 0, out, M.sh.S, M.sh.S, ex);
}

// comb table of the threshold key's V for b = V^r (thresholdkey.go:245) and V^Z (:306); Z is the widest exponent
int ensure_fix_v(pgpu_ctx* ctx) {
    if (ctx->fix_v.d) return PGPU_OK;
    if (ctx->tk_v.is_zero()) return fail(ctx, PGPU_ERR_STATE, "threshold key has no verification base V");
    return fixed_table_build(ctx, ctx->m_n2, ctx->tk_v, 32 * z_limbs(ctx), ctx->fix_v);
}

void protocols_free(pgpu_ctx* ctx) {
    program_free(ctx->prog_enc2); program_free(ctx->prog_rand); program_free(ctx->prog_alt1); program_free(ctx->prog_alt2);
    fixed_table_free(ctx->fix_h1); fixed_table_free(ctx->fix_h2); fixed_table_free(ctx->fix_v);
    if (ctx->d_rec2) { cudaFree(ctx->d_rec2); ctx->d_rec2 = nullptr; }
    dev_scrub_free(ctx->d_crt2, ctx->d_crt2_bytes); ctx->d_crt2 = nullptr; ctx->d_crt2_bytes = 0;
    program_free(ctx->prog_dec2_p); program_free(ctx->prog_dec2_q);
    program_free(ctx->prog_enc2q); program_free(ctx->prog_enc2p); program_free(ctx->prog_enc2f);
    ctx->enc2_crt_ready = false;
    modctx_free(ctx->m_p3); modctx_free(ctx->m_q3);
    ctx->crt2_ready = false;
}

// Public-key constants of the level-1 / level-2 shortcuts and the programs that need no secret.
int setup_level2(pgpu_ctx* ctx) {
    int rc;
    ModCtx &M1 = ctx->m_n, &M2 = ctx->m_n2, &M3 = ctx->m_n3;
    if ((rc = set_kconst(ctx, M2, K_NM, to_mont(M2, ctx->n)))) return rc;
    if ((rc = set_kconst(ctx, M2, K_NSM, to_mont(M2, ctx->n)))) return rc;
    if ((rc = set_kconst(ctx, M1, K_R4, (M1.R3 * M1.W1) % M1.N))) return rc;      // W^2 * R^2: third chunk of a record
    {   // Randomize with the r supplied: c * r^n mod n^2 (operations.go:67-69 = Add(ct, EncryptWithR(0, r)))
        Program& P = ctx->prog_rand;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_pow_shared(P, ctx->n, 0);
        P.emit(OP_MULI, 1); P.n_mul++;               // (r^n * R) * c * R^-1
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    if (!M3.ready) return PGPU_OK;
    BigU inv2;
    if (!BigU::modinv(BigU(2), M3.N, inv2)) return fail(ctx, PGPU_ERR_ARG, "n must be odd");
    if ((rc = set_kconst(ctx, M3, K_NEG1, to_mont(M3, M3.N - BigU(1))))) return rc;
    if ((rc = set_kconst(ctx, M3, K_C1, to_mont(M3, (ctx->n2 * inv2) % M3.N)))) return rc;
    if ((rc = set_kconst(ctx, M3, K_NM, to_mont(M3, ctx->n)))) return rc;
    if ((rc = set_kconst(ctx, M3, K_NSM, to_mont(M3, ctx->n2)))) return rc;
    {   // EncryptWithRAtLevel, level 2 (paillier.go:206-218): c = (1+n)^m * r^(n^2) mod n^3.  in0 = r, in1 = m
        Program& P = ctx->prog_enc2;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_pow_shared(P, ctx->n2, 0);
        const uint32_t RES = P.tbl_entries;
        P.emit(OP_STT, RES); P.use_slot(RES);
        P.emit(OP_LDI, 1);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_g_pow_level2(P, RES + 1);
        P.emit(OP_MULT, RES); P.n_mul++;
        P.emit(OP_MULC, K_ONE); P.n_mul++;
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    ctx->level2_ready = true;
    return PGPU_OK;
}

int encrypt2_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c) {
    if (!ctx->level2_ready) return fail(ctx, PGPU_ERR_UNSUPPORTED, "level 2: n^3 is wider than the built kernel shapes");
    const ModCtx& M = ctx->m_n3;
    const uint32_t wn = (uint32_t)ctx->wn;
    IoDesc ins[2] = {{r, wn, wn}, {m, 2 * wn, 2 * wn}};
    return run_vm(ctx, M, ctx->prog_enc2, count, ins, 2, c, M.sh.S, M.sh.S);
}

// Randomize (operations.go:67-69) with the r of the fresh encryption of zero supplied
int randomize_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, const uint32_t* r, uint32_t* out) {
    const ModCtx& M = ctx->m_n2;
    const uint32_t wn = (uint32_t)ctx->wn, S = M.sh.S;
    if (ctx->has_secret && ctx->has_enc_crt && !getenv("PGPU_NO_CRT_PROTOCOLS")) {
        // key holder: r^n = EncryptWithR(0, r) over p^2 and q^2 (encrypt_crt_dev), then one multiplication
        DEVBUF(zero, ctx, count * wn); DEVBUF(rn, ctx, count * S);
        CU(ctx, cudaMemsetAsync(zero.p, 0, count * wn * 4, ctx->stream));
        int rc = encrypt_crt_dev(ctx, count, zero.p, r, rn.p);
        if (rc) return rc;
        return modmul_dev(ctx, M, count, c, rn.p, out);
    }
    IoDesc ins[2] = {{r, wn, wn}, {c, S, S}};
    return run_vm(ctx, M, ctx->prog_rand, count, ins, 2, out, S, S);
}

// Decrypt at level 2 (paillier.go:292-340): c^lambda mod n^3, recoveryAlgorithm(s = 2), times lambda^-1 mod n^2.
// Secret constants for recover2_kernel; called from setup_crt once p, q are known.
int setup_level2_secret(pgpu_ctx* ctx) {
    if (!ctx->level2_ready) return PGPU_OK;
    const size_t h = ctx->wn, H = 2 * h;
    if (H > (size_t)CRT_MAXH) return PGPU_OK;      // level-2 decrypt unsupported at this width; reported on use
    const BigU lambda = (ctx->p - BigU(1)) * (ctx->q - BigU(1));
    const BigU RH = BigU::pow2(32 * H);
    const BigU& N2 = ctx->n2;
    BigU ninv, inv2, mu;
    if (!BigU::modinv(ctx->n, RH, ninv) || !BigU::modinv(BigU(2), N2, inv2)) return fail(ctx, PGPU_ERR_ARG, "n must be odd");
    if (!BigU::modinv(lambda % N2, N2, mu)) return fail(ctx, PGPU_ERR_NOT_INVERTIBLE, "lambda is not invertible mod n^2");
    const BigU R1 = RH % N2;
    std::vector<uint32_t> K;
    for (const BigU& x : {N2, ninv, (R1 * R1) % N2, (inv2 * R1) % N2, (mu * R1) % N2, (ctx->n * R1) % N2}) {
        auto l = x.limbs(H); K.insert(K.end(), l.begin(), l.end());
    }
    if (ctx->d_rec2) { cudaFree(ctx->d_rec2); ctx->d_rec2 = nullptr; }
    CU(ctx, cudaMalloc(&ctx->d_rec2, K.size() * 4));
    int rc;
    if ((rc = upload(ctx, ctx->d_rec2, K))) return rc;
    return setup_level2_crt(ctx);
}

// CRT constants of level-2 Decrypt (Crt2Params in aux.h).  Optional: without them decrypt2_dev takes the direct route.
static void crt2_prime_consts(std::vector<uint32_t>& K, const BigU& p, const BigU& q, size_t h) {
    const size_t H = 2 * h;
    const BigU p2 = p * p, R = BigU::pow2(32 * H), Rh = BigU::pow2(32 * h);
    BigU pinvH, cm, qinv, inv2;
    BigU::modinv(p, R, pinvH);
    BigU::modinv((q * (p - BigU(1))) % p2, p2, cm);
    BigU::modinv(q % p, p, qinv);
    BigU::modinv(BigU(2), p, inv2);
    auto push = [&](const BigU& x, size_t w) { auto l = x.limbs(w); K.insert(K.end(), l.begin(), l.end()); };
    push(p2, H); push(pinvH, H); push((cm * (R % p2)) % p2, H);
    const BigU rh = Rh % p;
    push(p, h); push((rh * rh) % p, h); push((((rh * rh) % p) * rh) % p, h);
    push((qinv * rh) % p, h); push(((((inv2 * q) % p) * q) % p * rh) % p, h); push(rh, h);
}

// EncryptWithRAtLevel(level 2) for the holder of p, q (paillier.go:206-218; SecretKey embeds PublicKey, :29-34):
// r^(n^2) mod n^3 from r^e mod q^3 and mod p^3 with e = (n^2 mod phi) + phi, phi = P^2 (P-1) of that prime --
// for r coprime to P that is r^(n^2) by Euler, for P | r both sides are 0 since e >= 3 -- then Garner's step and the
// g^m factor over n^3.  Exponents of ~3k bits over 3k-bit moduli instead of 4096 bits over 6144.
static int setup_encrypt2_crt(pgpu_ctx* ctx) {
    ctx->enc2_crt_ready = false;
    ModCtx &P3 = ctx->m_p3, &Q3 = ctx->m_q3, &N3 = ctx->m_n3;
    if (!ctx->level2_ready || ctx->wn > (size_t)P3.sh.S || P3.sh.S > N3.sh.S) return PGPU_OK;
    const BigU &p = ctx->p, &q = ctx->q;
    const BigU phip = p * p * (p - BigU(1)), phiq = q * q * (q - BigU(1));
    const BigU ep = (ctx->n2 % phip) + phip, eq = (ctx->n2 % phiq) + phiq;
    BigU q3inv;
    if (!BigU::modinv(Q3.N % P3.N, P3.N, q3inv)) return fail(ctx, PGPU_ERR_ARG, "p and q are not coprime");
    int rc;
    if ((rc = set_kconst(ctx, P3, K_CRT, q3inv))) return rc;
    if ((rc = set_kconst(ctx, N3, K_CRT, (Q3.N * N3.R2) % N3.N))) return rc;
    {   // x_q = r^eq mod q^3.  in0 = r
        Program& P = ctx->prog_enc2q;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_pow_shared(P, eq, 0);
        P.emit(OP_MULC, K_ONE); P.n_mul++;
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    {   // t = (r^ep - x_q) * q^-3 mod p^3.  in0 = r, in1 = x_q
        Program& P = ctx->prog_enc2p;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_pow_shared(P, ep, 0);
        const uint32_t XP = P.tbl_entries, XQ = XP + 1;
        P.emit(OP_STT, XP); P.use_slot(XP);
        P.emit(OP_LDI, 1);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        P.emit(OP_STT, XQ); P.use_slot(XQ);
        P.emit(OP_LDT, XP);
        P.emit(OP_SUBT, XQ);
        P.emit(OP_MULC, K_CRT); P.n_mul++;      // plain constant: leaves Montgomery form
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    {   // c = (x_q + q^3*t) * (1+n)^m mod n^3.  in0 = t, in1 = x_q, in2 = m
        Program& P = ctx->prog_enc2f;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_CRT); P.n_mul++;
        P.emit(OP_STT, 0); P.use_slot(0);
        P.emit(OP_LDI, 1);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        P.emit(OP_ADDT, 0);                     // r^(n^2) mod n^3 (x_q + q^3*t < n^3)
        P.emit(OP_STT, 0);
        P.emit(OP_LDI, 2);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_g_pow_level2(P, 1);
        P.emit(OP_MULT, 0); P.n_mul++;
        P.emit(OP_MULC, K_ONE); P.n_mul++;
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    ctx->enc2_crt_ready = true;
    return PGPU_OK;
}

int setup_level2_crt(pgpu_ctx* ctx) {
    ctx->crt2_ready = false;
    const BigU &p = ctx->p, &q = ctx->q;
    const BigU p3 = p * p * p, q3 = q * q * q;
    Shape s1, s2;
    if (!pick_shape(p3.v.size(), s1) || !pick_shape(q3.v.size(), s2) || s1.S != s2.S) return PGPU_OK;
    const size_t h = ctx->crt_h, H = 2 * h;
    const uint32_t S3 = ctx->m_n3.sh.S, Sp = s1.S;
    if (H > (size_t)CRT_MAXH || H > Sp || S3 > 2 * Sp || p.v.size() > h || q.v.size() > h || 2 * H < 2 * ctx->wn) return PGPU_OK;
    int rc;
    modctx_free(ctx->m_p3); modctx_free(ctx->m_q3);
    if ((rc = modctx_init(ctx, ctx->m_p3, p3))) return rc;
    if ((rc = modctx_init(ctx, ctx->m_q3, q3))) return rc;
    if ((rc = build_decrypt_half(ctx, ctx->prog_dec2_p, p - BigU(1)))) return rc;
    if ((rc = build_decrypt_half(ctx, ctx->prog_dec2_q, q - BigU(1)))) return rc;
    std::vector<uint32_t> K;
    crt2_prime_consts(K, p, q, h);
    ctx->crt2_cq_off = K.size();
    crt2_prime_consts(K, q, p, h);
    ctx->crt2_g_off = K.size();
    const BigU p2 = p * p, q2 = q * q, R = BigU::pow2(32 * H);
    BigU q2inv;
    if (!BigU::modinv(q2 % p2, p2, q2inv)) return fail(ctx, PGPU_ERR_ARG, "p and q are not coprime");
    for (const BigU& x : {(q2inv * (R % p2)) % p2, q2}) { auto l = x.limbs(H); K.insert(K.end(), l.begin(), l.end()); }
    dev_scrub_free(ctx->d_crt2, ctx->d_crt2_bytes); ctx->d_crt2 = nullptr; ctx->d_crt2_bytes = 0;
    CU(ctx, cudaMalloc(&ctx->d_crt2, K.size() * 4));
    ctx->d_crt2_bytes = K.size() * 4;
    if ((rc = upload(ctx, ctx->d_crt2, K))) return rc;
    ctx->crt2_np0[0] = mont_np0(p.v[0]); ctx->crt2_np0[1] = mont_np0(p2.v[0]);
    ctx->crt2_np0[2] = mont_np0(q.v[0]); ctx->crt2_np0[3] = mont_np0(q2.v[0]);
    ctx->crt2_ready = true;
    return setup_encrypt2_crt(ctx);
}

int encrypt2_crt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c) {
    if (!ctx->has_secret) return fail(ctx, PGPU_ERR_STATE, "EncryptWithRAtLevel (secret key): no secret key loaded");
    if (!ctx->enc2_crt_ready) return encrypt2_dev(ctx, count, m, r, c);
    const ModCtx &P3 = ctx->m_p3, &Q3 = ctx->m_q3, &N3 = ctx->m_n3;
    const uint32_t Sp = P3.sh.S, wn = (uint32_t)ctx->wn;
    DEVBUF(xq, ctx, count * Sp); DEVBUF(t, ctx, count * Sp);
    int rc;
    IoDesc iq[1] = {{r, wn, wn}};
    if ((rc = run_vm(ctx, Q3, ctx->prog_enc2q, count, iq, 1, xq.p, Sp, Sp))) return rc;
    IoDesc ip[2] = {{r, wn, wn}, {xq.p, Sp, Sp}};
    if ((rc = run_vm(ctx, P3, ctx->prog_enc2p, count, ip, 2, t.p, Sp, Sp))) return rc;
    IoDesc fin[3] = {{t.p, Sp, Sp}, {xq.p, Sp, Sp}, {m, 2 * wn, 2 * wn}};
    return run_vm(ctx, N3, ctx->prog_enc2f, count, fin, 3, c, N3.sh.S, N3.sh.S);
}

static int decrypt2_crt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, uint32_t* m) {
    const ModCtx &P3 = ctx->m_p3, &Q3 = ctx->m_q3;
    const uint32_t S3 = ctx->m_n3.sh.S, Sp = P3.sh.S;
    const uint32_t hi_limbs = S3 > Sp ? S3 - Sp : 0;
    DEVBUF(xp, ctx, count * Sp); DEVBUF(xq, ctx, count * Sp);
    DEVBUF(mhp, ctx, count * 2 * ctx->crt_h); DEVBUF(mhq, ctx, count * 2 * ctx->crt_h);
    int rc;
    IoDesc ins[2] = {{c, S3, std::min(S3, Sp)}, {c + Sp, S3, hi_limbs}};
    if ((rc = run_vm(ctx, P3, ctx->prog_dec2_p, count, ins, 2, xp.p, Sp, Sp))) return rc;          // c^(p-1) mod p^3
    if ((rc = run_vm(ctx, Q3, ctx->prog_dec2_q, count, ins, 2, xq.p, Sp, Sp))) return rc;          // c^(q-1) mod q^3
    Crt2Params C{};
    C.n_items = (uint32_t)count; C.h = ctx->crt_h;
    C.cp = ctx->d_crt2; C.cq = ctx->d_crt2 + ctx->crt2_cq_off; C.garner = ctx->d_crt2 + ctx->crt2_g_off;
    C.np0_p = ctx->crt2_np0[0]; C.np0_p2 = ctx->crt2_np0[1]; C.np0_q = ctx->crt2_np0[2]; C.np0_q2 = ctx->crt2_np0[3];
    C.xp = xp.p; C.xq = xq.p; C.x_stride = Sp;
    C.out = m; C.out_limbs = 2 * (uint32_t)ctx->wn;
    C.mp = mhp.p; C.mq = mhq.p;
    CU(ctx, crt2_launch(C, ctx->stream));
    ctx->launches += 3;
    return PGPU_OK;
}

int decrypt2_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, uint32_t* m) {
    if (!ctx->has_secret) return fail(ctx, PGPU_ERR_STATE, "Decrypt: no secret key loaded");
    if (!ctx->level2_ready || !ctx->d_rec2) return fail(ctx, PGPU_ERR_UNSUPPORTED, "level-2 Decrypt is not available at this key size");
    if (ctx->crt2_ready && !getenv("PGPU_NO_CRT2")) return decrypt2_crt_dev(ctx, count, c, m);
    const ModCtx& M = ctx->m_n3;
    const uint32_t S = M.sh.S;
    const BigU lambda = (ctx->p - BigU(1)) * (ctx->q - BigU(1));
    DEVBUF(tmp, ctx, count * S);
    int rc;
    if ((rc = modexp_shared_dev(ctx, M, count, c, lambda, tmp.p))) return rc;          // :296
    const uint32_t H = 2 * (uint32_t)ctx->wn;
    Recover2Params R{(uint32_t)count, (int)ctx->wn, ctx->d_rec2, mont_np0(ctx->n2.v[0]), tmp.p, S, S, m, H};
    CU(ctx, recover2_launch(R, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

// AltEncryptWithRAtLevel (paillier.go:221-238): c = (1+n)^m * h_s^(r mod K) mod n^(s+1), h_s a per-key
// fixed base -> comb table, no squarings.  r: n-width records of which the low kbits bits are used (r mod K).
int setup_alt(pgpu_ctx* ctx) {
    int rc;
    const BigU& H = ctx->alt_h;
    {
        ModCtx& M = ctx->m_n2;
        const BigU h1 = BigU::modexp((ctx->n - (H % ctx->n)) % ctx->n, ctx->n, ctx->n2);     // :420-423 (H < n)
        if ((rc = fixed_table_build(ctx, M, h1, ctx->alt_kbits, ctx->fix_h1))) return rc;
        Program& P = ctx->prog_alt1;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_g_pow_level1(P);
        P.emit(OP_STT, 0); P.use_slot(0);
        emit_pow_fixed(P, ctx->fix_h1);
        P.emit(OP_MULT, 0); P.n_mul++;
        P.emit(OP_MULC, K_ONE); P.n_mul++;
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    if (ctx->level2_ready) {
        ModCtx& M = ctx->m_n3;
        const BigU h2 = BigU::modexp(ctx->n2 - H, ctx->n2, ctx->n3);                         // :428-431
        if ((rc = fixed_table_build(ctx, M, h2, ctx->alt_kbits, ctx->fix_h2))) return rc;
        Program& P = ctx->prog_alt2;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_g_pow_level2(P, 1);
        P.emit(OP_STT, 0); P.use_slot(0);
        emit_pow_fixed(P, ctx->fix_h2);
        P.emit(OP_MULT, 0); P.n_mul++;
        P.emit(OP_MULC, K_ONE); P.n_mul++;
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    ctx->has_alt = true;
    return PGPU_OK;
}

int alt_encrypt_dev(pgpu_ctx* ctx, int level, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c) {
    if (!ctx->has_alt) return fail(ctx, PGPU_ERR_STATE, "AltEncrypt: H and K were not loaded (pgpu_ctx_set_alt_generator)");
    const uint32_t wn = (uint32_t)ctx->wn;
    if (level == 1) {
        const ModCtx& M = ctx->m_n2;
        IoDesc ins[1] = {{m, wn, wn}};
        return run_vm(ctx, M, ctx->prog_alt1, count, ins, 1, c, M.sh.S, M.sh.S, ExpDesc{r, wn, ctx->alt_kbits, ctx->fix_h1.d});
    }
    if (level == 2) {
        if (!ctx->level2_ready) return fail(ctx, PGPU_ERR_UNSUPPORTED, "level 2: n^3 is wider than the built kernel shapes");
        const ModCtx& M = ctx->m_n3;
        IoDesc ins[1] = {{m, 2 * wn, 2 * wn}};
        return run_vm(ctx, M, ctx->prog_alt2, count, ins, 1, c, M.sh.S, M.sh.S, ExpDesc{r, wn, ctx->alt_kbits, ctx->fix_h2.d});
    }
    return fail(ctx, PGPU_ERR_ARG, "encryption level must be 1 or 2");
}

// ExtractRandonness (operations.go:75-91): z = (g^v)^-1 * c mod n^(s+1) with v = Decrypt(c), then
// z^((n^s)^-1 mod lambda) mod n.  (g^v)^-1 = g^(n^s - v) because g = 1+n has order n^s mod n^(s+1).
int extract_randomness_dev(pgpu_ctx* ctx, int level, size_t count, const uint32_t* c, uint32_t* out) {
    if (!ctx->has_secret) return fail(ctx, PGPU_ERR_STATE, "ExtractRandonness: no secret key loaded");
    if (level != 1 && level != 2) return fail(ctx, PGPU_ERR_ARG, "encryption level must be 1 or 2");
    if (level == 2 && !ctx->level2_ready) return fail(ctx, PGPU_ERR_UNSUPPORTED, "level 2: n^3 is wider than the built kernel shapes");
    const ModCtx& M = level == 1 ? ctx->m_n2 : ctx->m_n3;
    const ModCtx& MN = ctx->m_n;
    const uint32_t S = M.sh.S, Sn = MN.sh.S, wn = (uint32_t)ctx->wn, wv = level == 1 ? wn : 2 * wn;
    const BigU lambda = (ctx->p - BigU(1)) * (ctx->q - BigU(1));
    const BigU& ns = level == 1 ? ctx->n : ctx->n2;
    BigU ns_inv;
    if (!BigU::modinv(ns % lambda, lambda, ns_inv)) return fail(ctx, PGPU_ERR_NOT_INVERTIBLE, "n^s is not invertible mod lambda");   // :79
    int rc;
    DEVBUF(v, ctx, count * wv); DEVBUF(z, ctx, count * S);
    if ((rc = level == 1 ? decrypt_dev(ctx, count, c, v.p) : decrypt2_dev(ctx, count, c, v.p))) return rc;     // :81
    {   // z = g^(n^s - v) * c mod n^(s+1)   (:82-86)
        const std::string key = "xr-z:" + std::to_string(level);
        Program* P = cached_program(ctx, key);
        if (!P) {
            Program np;
            np.emit(OP_LDI, 0);
            np.emit(OP_MULC, K_R2); np.n_mul++;
            np.emit(OP_STT, 0); np.use_slot(0);
            np.emit(OP_LDC, K_NSM);
            np.emit(OP_SUBT, 0);                     // n^s - v
            if (level == 1) emit_g_pow_level1(np); else emit_g_pow_level2(np, 1);
            np.emit(OP_MULI, 1); np.n_mul++;         // * c, leaves Montgomery form
            np.emit(OP_STO, 0);
            if ((rc = program_upload(ctx, np))) return rc;
            P = &(ctx->prog_cache[key] = np);
        }
        IoDesc ins[2] = {{v.p, wv, wv}, {c, S, S}};
        if ((rc = run_vm(ctx, M, *P, count, ins, 2, z.p, S, S))) return rc;
    }
    {   // (z mod n)^(ns_inv) mod n   (:88); z = sum chunk_k * 2^(32*Sn*k)
        const uint32_t chunks = (S + Sn - 1) / Sn;
        if (chunks > 3) return fail(ctx, PGPU_ERR_UNSUPPORTED, "ExtractRandonness: ciphertext record too wide for the n-sized kernel");
        const std::string key = "xr-p:" + std::to_string(level) + ":" + ns_inv.hex();
        Program* P = cached_program(ctx, key);
        if (!P) {
            Program np;
            static const uint32_t kslot[3] = {K_R2, K_R3, K_R4};
            for (uint32_t k = 0; k < chunks; ++k) {
                np.emit(OP_LDI, k);
                np.emit(OP_MULC, kslot[k]); np.n_mul++;       // chunk_k * R^k, Montgomery form
                if (k) np.emit(OP_ADDT, 0);
                if (k + 1 < chunks) { np.emit(OP_STT, 0); np.use_slot(0); }
            }
            emit_pow_shared(np, ns_inv, 0);
            np.emit(OP_MULC, K_ONE); np.n_mul++;
            np.emit(OP_STO, 0);
            if ((rc = program_upload(ctx, np))) return rc;
            P = &(ctx->prog_cache[key] = np);
        }
        IoDesc ins[3];
        for (uint32_t k = 0; k < chunks; ++k) ins[k] = IoDesc{z.p + k * Sn, S, std::min(Sn, S - k * Sn)};
        if ((rc = run_vm(ctx, MN, *P, count, ins, (int)chunks, out, wn, wn))) return rc;
    }
    return PGPU_OK;
}

// out[i] = b1[i]^e1[i] * b2[i]^e2 mod M: per-item exponents e1 (exp.bits bits, fixed window) and one shared exponent e2,
// interleaved so that the squarings are shared (Straus / Shamir): per window w squarings, one multiplication by
// T1[bits of e1] (OP_WIN) and, where e2's window is non-zero, one by T2[bits of e2] (the host knows e2: a plain OP_MULT).
static int modexp_dual_io(pgpu_ctx* ctx, const ModCtx& M, size_t count, const IoDesc& b1, const ExpDesc& exp, const IoDesc& b2,
                          const BigU& e2, const std::string& e2_name, uint32_t* out) {
    const size_t bits = std::max<size_t>(exp.bits, e2.bitlen());
    const std::string key = "dual:" + std::to_string(M.sh.S) + ":" + std::to_string(exp.bits) + ":" + e2_name;
    Program* P = cached_program(ctx, key);
    if (!P) {
        Program np;
        int w = 1;
        { double best = 1e300; for (int c = 1; c <= 6; ++c) { double cost = 2.0 * (1u << c) + 2.0 * (double)bits / c; if (cost < best) { best = cost; w = c; } } }
        const uint32_t n = 1u << w, B = n;                    // T1 = slots [0, n), T2 = slots [n, 2n)
        np.emit(OP_LDI, 1);
        np.emit(OP_MULC, K_R2); np.n_mul++;
        np.emit(OP_STT, B + 1);
        for (uint32_t i = 2; i < n; ++i) { np.emit(OP_MULT, B + 1); np.n_mul++; np.emit(OP_STT, B + i); }
        np.use_slot(B + n - 1);
        np.emit(OP_LDI, 0);
        np.emit(OP_MULC, K_R2); np.n_mul++;
        np.emit(OP_STT, 1);
        for (uint32_t i = 2; i < n; ++i) { np.emit(OP_MULT, 1); np.n_mul++; np.emit(OP_STT, i); }
        np.emit(OP_LDC, K_R1); np.emit(OP_STT, 0);
        const size_t nwin = (bits + w - 1) / w;
        for (size_t k = nwin; k-- > 0;) {
            np.emit(OP_WIN, (uint32_t)(k * w) | ((uint32_t)w << 20));
            np.n_sqr += w; np.n_mul++;
            uint32_t val = 0;
            for (int j = w - 1; j >= 0; --j) val = (val << 1) | (e2.bit(k * w + j) ? 1u : 0u);
            if (val) { np.emit(OP_MULT, B + val); np.n_mul++; }
        }
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STO, 0);
        int rc = program_upload(ctx, np);
        if (rc) return rc;
        P = &(ctx->prog_cache[key] = np);
    }
    IoDesc ins[2] = {b1, b2};
    return run_vm(ctx, M, *P, count, ins, 2, out, M.sh.S, M.sh.S, exp);
}

// base[i]^exp[i] mod n^3 for the holder of p, q: the same exponentiation over q^3 and p^3 (half-width moduli, the
// exponent unreduced so that no assumption on the base is needed), Garner's step in the tail of the p^3 program and
// x = x_q + q^3*t over n^3.  Bit-identical to modexp_items_io(M3, ...) -- both return the canonical residue.
static int modexp3_items_crt(pgpu_ctx* ctx, size_t count, const IoDesc& base, const ExpDesc& exp, uint32_t* out) {
    const ModCtx &P3 = ctx->m_p3, &Q3 = ctx->m_q3, &N3 = ctx->m_n3;
    const uint32_t Sp = P3.sh.S, S3 = N3.sh.S;
    int rc;
    auto prefix = [](Program& np) {             // (lo + hi*R) mod P^3 in Montgomery form, R = 2^(32*Sp)
        np.emit(OP_LDI, 0);
        np.emit(OP_MULC, K_R2); np.n_mul++;
        np.emit(OP_STT, 0); np.use_slot(0);
        np.emit(OP_LDI, 1);
        np.emit(OP_MULC, K_R3); np.n_mul++;
        np.emit(OP_ADDT, 0);
    };
    const std::string kq = "crt3q:" + std::to_string(exp.bits), kp = "crt3p:" + std::to_string(exp.bits), kf = "crt3f";
    Program* Pq = cached_program(ctx, kq);
    if (!Pq) {
        Program np;
        prefix(np);
        emit_pow_items(np, exp.bits, 0);
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, np))) return rc;
        Pq = &(ctx->prog_cache[kq] = np);
    }
    Program* Pp = cached_program(ctx, kp);
    if (!Pp) {
        Program np;
        prefix(np);
        emit_pow_items(np, exp.bits, 0);
        const uint32_t XP = np.tbl_entries, XQ = XP + 1;
        np.emit(OP_STT, XP); np.use_slot(XP);
        np.emit(OP_LDI, 2);
        np.emit(OP_MULC, K_R2); np.n_mul++;
        np.emit(OP_STT, XQ); np.use_slot(XQ);
        np.emit(OP_LDT, XP);
        np.emit(OP_SUBT, XQ);
        np.emit(OP_MULC, K_CRT); np.n_mul++;
        np.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, np))) return rc;
        Pp = &(ctx->prog_cache[kp] = np);
    }
    Program* Pf = cached_program(ctx, kf);
    if (!Pf) {
        Program np;
        np.emit(OP_LDI, 0);
        np.emit(OP_MULC, K_CRT); np.n_mul++;    // q^3 * t, Montgomery form
        np.emit(OP_STT, 0); np.use_slot(0);
        np.emit(OP_LDI, 1);
        np.emit(OP_MULC, K_R2); np.n_mul++;
        np.emit(OP_ADDT, 0);
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, np))) return rc;
        Pf = &(ctx->prog_cache[kf] = np);
    }
    DEVBUF(xq, ctx, count * Sp); DEVBUF(t, ctx, count * Sp);
    const uint32_t lo = std::min(base.limbs, Sp), hi = base.limbs > Sp ? base.limbs - Sp : 0;
    IoDesc iq[2] = {{base.ptr, base.stride, lo, base.div}, {base.ptr + Sp, base.stride, hi, base.div}};
    if ((rc = run_vm(ctx, Q3, *Pq, count, iq, 2, xq.p, Sp, Sp, exp))) return rc;
    IoDesc ip[3] = {iq[0], iq[1], {xq.p, Sp, Sp}};
    if ((rc = run_vm(ctx, P3, *Pp, count, ip, 3, t.p, Sp, Sp, exp))) return rc;
    IoDesc fin[2] = {{t.p, Sp, Sp}, {xq.p, Sp, Sp}};
    return run_vm(ctx, N3, *Pf, count, fin, 2, out, S3, S3);
}

// Exponentiations of the protocols that the holder of p, q may run over the prime powers (same results, about half the work)
struct KeyHolderPow {
    pgpu_ctx* ctx; bool crt; DevBuf zero; uint32_t wn, S2, e2bits;
    KeyHolderPow(pgpu_ctx* c, size_t max_items)
        : ctx(c), crt(c->has_secret && c->enc2_crt_ready && c->has_enc_crt && !getenv("PGPU_NO_CRT_PROTOCOLS")),
          zero(c, crt ? max_items * 2 * c->wn : 1), wn((uint32_t)c->wn), S2(c->m_n2.sh.S), e2bits((uint32_t)c->n2.bitlen()) {
        if (crt && zero.p) cudaMemsetAsync(zero.p, 0, max_items * 2 * c->wn * 4, c->stream);
    }
    // base^e mod n^3 with per-item exponents e (n^2-wide records)
    int pow3_items(size_t k, const IoDesc& base, const uint32_t* e_ptr, uint32_t* out) {
        const ExpDesc ed{e_ptr, S2, e2bits, nullptr};
        return crt ? modexp3_items_crt(ctx, k, base, ed, out) : modexp_items_io(ctx, ctx->m_n3, k, base, ed, out);
    }
    // r^(n^2) mod n^3 = EncryptWithRAtLevel(0, r); r: n-wide records
    int pow3_n2(size_t k, const uint32_t* r, uint32_t* out) {
        return crt ? encrypt2_crt_dev(ctx, k, zero.p, r, out) : modexp_shared_io(ctx, ctx->m_n3, k, IoDesc{r, wn, wn}, ctx->n2, out);
    }
    // r^n mod n^2 = EncryptWithR(0, r)
    int pow2_n(size_t k, const uint32_t* r, uint32_t* out) {
        return crt ? encrypt_crt_dev(ctx, k, zero.p, r, out) : modexp_shared_io(ctx, ctx->m_n2, k, IoDesc{r, wn, wn}, ctx->n, out);
    }
};

// NestedRandomize (operations.go:96-118) with a, b supplied: ct^(a^n mod n^2) * b^(n^2) mod n^3
int nested_randomize_dev(pgpu_ctx* ctx, size_t count, const uint32_t* ct, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    if (!ctx->level2_ready) return fail(ctx, PGPU_ERR_UNSUPPORTED, "level 2: n^3 is wider than the built kernel shapes");
    const ModCtx &M2 = ctx->m_n2, &M3 = ctx->m_n3;
    const uint32_t S2 = M2.sh.S, S3 = M3.sh.S;
    DEVBUF(an, ctx, count * S2); DEVBUF(bn2, ctx, count * S3); DEVBUF(t, ctx, count * S3);
    int rc;
    KeyHolderPow kh(ctx, count);                 // over p^3, q^3 when the context holds the secret key
    if (kh.zero.err != cudaSuccess) return fail(ctx, PGPU_ERR_CUDA, "device allocation failed");
    if ((rc = kh.pow2_n(count, a, an.p))) return rc;                                                  // a^n mod n^2      :108
    if (!kh.crt && !getenv("PGPU_NO_DUAL_EXP"))                                                       // ct^(a^n) * b^(n^2), squarings shared :109-114
        return modexp_dual_io(ctx, M3, count, IoDesc{ct, S3, S3}, ExpDesc{an.p, S2, kh.e2bits, nullptr}, IoDesc{b, kh.wn, kh.wn}, ctx->n2, "n2", out);
    if ((rc = kh.pow3_n2(count, b, bn2.p))) return rc;                                                // b^(n^2) mod n^3  :109
    if ((rc = kh.pow3_items(count, IoDesc{ct, S3, S3}, an.p, t.p))) return rc;                        // ct^(a^n)         :112
    return modmul_dev(ctx, M3, count, t.p, bn2.p, out);                                               // :113-114
}

// digest of RandomOracleBit(ct1, ct2, x, y, alpha): ct1 is skipped (random_oracle.go:24-26)
static int ddleq_hash(pgpu_ctx* ctx, size_t total, uint32_t secpar, const uint32_t* ct2, const uint32_t* x, const uint32_t* y,
                      const uint32_t* alpha, uint32_t* digest) {
    const uint32_t S3 = ctx->m_n3.sh.S, wn = (uint32_t)ctx->wn;
    const uint32_t* seg[4] = {ct2, x, y, alpha};
    const uint32_t stride[4] = {S3, wn, wn, S3};
    const int limbs[4] = {(int)S3, (int)wn, (int)wn, (int)S3};
    const uint32_t div[4] = {secpar, 1, 1, 1};
    return sha_dev(ctx, total, 4, seg, stride, limbs, digest, div);
}

// ProveDDLEQ (ddleq.go:27-127) for `count` statements (ct1, ct2, a, b) with secpar instances each; the
// x, y in Z*_n of every instance are supplied (ddleq.go:71-79 draws them).  The values that do not
// depend on the instance (sanity check :62-69, ExtractRandonness :103, a^n :104, a^-1 :96) are computed
// once per statement.  *d_bad = first statement whose sanity check fails, else 0xffffffff.
int ddleq_prove_dev(pgpu_ctx* ctx, size_t count, uint32_t secpar, const uint32_t* ct1, const uint32_t* ct2, const uint32_t* a, const uint32_t* b,
                    const uint32_t* x, const uint32_t* y, uint32_t* alpha, uint32_t* e, uint32_t* f, uint32_t* d_bad) {
    if (!ctx->has_secret) return fail(ctx, PGPU_ERR_STATE, "ProveDDLEQ: no secret key loaded");
    if (!ctx->level2_ready) return fail(ctx, PGPU_ERR_UNSUPPORTED, "level 2: n^3 is wider than the built kernel shapes");
    if (secpar == 0) return fail(ctx, PGPU_ERR_ARG, "ProveDDLEQ: secpar must be positive");
    const ModCtx &M2 = ctx->m_n2, &M3 = ctx->m_n3;
    const uint32_t S2 = M2.sh.S, S3 = M3.sh.S, wn = (uint32_t)ctx->wn, e2bits = (uint32_t)ctx->n2.bitlen();
    const size_t total = count * secpar;
    int rc;
    KeyHolderPow kh(ctx, std::max(total, count));   // the prover holds p, q: exponentiations run over the prime powers
    if (kh.zero.err != cudaSuccess) return fail(ctx, PGPU_ERR_CUDA, "device allocation failed");
    auto pow3_items = [&](size_t k, const IoDesc& base, const uint32_t* e_ptr, uint32_t* out) { return kh.pow3_items(k, base, e_ptr, out); };
    auto pow3_n2 = [&](size_t k, const uint32_t* r, uint32_t* out) { return kh.pow3_n2(k, r, out); };
    auto pow2_n = [&](size_t k, const uint32_t* r, uint32_t* out) { return kh.pow2_n(k, r, out); };
    // ---- per statement
    DEVBUF(an, ctx, count * S2); DEVBUF(bn2, ctx, count * S3); DEVBUF(t3, ctx, count * S3); DEVBUF(san, ctx, count * S3);
    DEVBUF(s, ctx, count * wn); DEVBUF(c0, ctx, count * S3); DEVBUF(ainv, ctx, count * S2); DEVBUF(a2, ctx, count * S2);
    DEVBUF(flags, ctx, (count + 3) / 4 + 1);
    if ((rc = pow2_n(count, a, an.p))) return rc;                                                                            // a^n        :63,104
    if ((rc = pow3_n2(count, b, bn2.p))) return rc;                                                                          // b^(n^2)    :64
    if ((rc = pow3_items(count, IoDesc{ct1, S3, S3}, an.p, t3.p))) return rc;                                                // ct1^(a^n)  :63
    if ((rc = modmul_dev(ctx, M3, count, t3.p, bn2.p, san.p))) return rc;                                                    // :64-65
    CU(ctx, equal_launch(san.p, ct2, S3, (uint32_t)count, (uint8_t*)flags.p, ctx->stream));                                  // :67
    ctx->launches++;
    CU(ctx, first_zero_launch((const uint8_t*)flags.p, (uint32_t)count, d_bad, ctx->stream));
    ctx->launches++;
    if ((rc = extract_randomness_dev(ctx, 2, count, ct1, s.p))) return rc;                                                   // s          :103
    if ((rc = pow3_items(count, IoDesc{s.p, wn, wn}, an.p, t3.p))) return rc;                                                // s^(a^n)    :107
    if ((rc = modmul_io(ctx, M3, count, IoDesc{t3.p, S3, S3}, IoDesc{b, wn, wn}, c0.p))) return rc;                          // * b        :108
    CU(ctx, resize_launch(a, wn, wn, a2.p, S2, (uint32_t)count, ctx->stream));
    ctx->launches++;
    DEVBUF(badinv, ctx, 1);
    if ((rc = modinv_batch_dev(ctx, M2, count, a2.p, ainv.p, badinv.p))) return rc;                                                // a^-1 mod n^2 :96
    // ---- per instance
    DEVBUF(xn, ctx, total * S2); DEVBUF(yn2, ctx, total * S3); DEVBUF(u3, ctx, total * S3); DEVBUF(dig, ctx, total * 8);
    if ((rc = pow2_n(total, x, xn.p))) return rc;                                                                            // x^n        :81
    if ((rc = pow3_n2(total, y, yn2.p))) return rc;                                                                          // y^(n^2)    :82
    if ((rc = pow3_items(total, IoDesc{ct1, S3, S3, secpar}, xn.p, u3.p))) return rc;                                        // ct1^xn     :85
    if ((rc = modmul_dev(ctx, M3, total, u3.p, yn2.p, alpha))) return rc;                                                    // alpha      :86-87
    if ((rc = ddleq_hash(ctx, total, secpar, ct2, x, y, alpha, dig.p))) return rc;                                           // challenge  :91
    // (e, f) = (x, y) where the challenge bit is 0 ...
    CU(ctx, resize_launch(x, wn, wn, e, S2, (uint32_t)total, ctx->stream));
    CU(ctx, resize_launch(y, wn, wn, f, S3, (uint32_t)total, ctx->stream));
    ctx->launches += 2;
    // ... and only the instances whose bit is 1 pay for the second half (:94-114): compact them
    std::vector<uint32_t> hd(total * 8), idx;
    CU(ctx, cudaMemcpyAsync(hd.data(), dig.p, total * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < total; ++i) if (hd[i * 8] & 1u) idx.push_back((uint32_t)i);
    const size_t n1 = idx.size();
    if (n1 == 0) return PGPU_OK;
    DEVBUF(didx, ctx, n1); DEVBUF(xg, ctx, n1 * wn); DEVBUF(yg, ctx, n1 * wn); DEVBUF(xng, ctx, n1 * S2); DEVBUF(ainvg, ctx, n1 * S2);
    DEVBUF(c0g, ctx, n1 * S3); DEVBUF(sg, ctx, n1 * wn);
    DEVBUF(e1, ctx, n1 * S2); DEVBUF(en, ctx, n1 * S2); DEVBUF(v3, ctx, n1 * S3); DEVBUF(f1, ctx, n1 * S3);
    if ((rc = upload(ctx, didx.p, idx))) return rc;
    const uint32_t N1 = (uint32_t)n1;
    CU(ctx, gather_launch(x, wn, didx.p, 1, xg.p, N1, ctx->stream));
    CU(ctx, gather_launch(y, wn, didx.p, 1, yg.p, N1, ctx->stream));
    CU(ctx, gather_launch(xn.p, S2, didx.p, 1, xng.p, N1, ctx->stream));
    CU(ctx, gather_launch(ainv.p, S2, didx.p, secpar, ainvg.p, N1, ctx->stream));
    CU(ctx, gather_launch(c0.p, S3, didx.p, secpar, c0g.p, N1, ctx->stream));
    CU(ctx, gather_launch(s.p, wn, didx.p, secpar, sg.p, N1, ctx->stream));
    ctx->launches += 6;
    if ((rc = modmul_io(ctx, M2, n1, IoDesc{xg.p, wn, wn}, IoDesc{ainvg.p, S2, S2}, e1.p))) return rc;                       // e = x*a^-1 mod n^2 :94-99
    if ((rc = modexp_shared_dev(ctx, M2, n1, e1.p, ctx->n, en.p))) return rc;                                                // e^n        :105
    if ((rc = pow3_items(n1, IoDesc{c0g.p, S3, S3}, en.p, u3.p))) return rc;                                                 // (s^an*b)^en :109
    if ((rc = modinv_batch_dev(ctx, M3, n1, u3.p, v3.p, badinv.p))) return rc;                                               // ^-1        :110
    if ((rc = pow3_items(n1, IoDesc{sg.p, wn, wn}, xng.p, u3.p))) return rc;                                                 // s^xn       :112
    if ((rc = modmul_dev(ctx, M3, n1, v3.p, u3.p, v3.p))) return rc;                                                         // c          :112
    if ((rc = modmul_io(ctx, M3, n1, IoDesc{yg.p, wn, wn}, IoDesc{v3.p, S3, S3}, f1.p))) return rc;                          // f = y*c    :113-114
    CU(ctx, scatter_launch(e1.p, S2, didx.p, e, N1, ctx->stream));
    CU(ctx, scatter_launch(f1.p, S3, didx.p, f, N1, ctx->stream));
    ctx->launches += 2;
    return PGPU_OK;
}

// VerifyDDLEQProof (ddleq.go:44-53,129-153): ok[i] for every instance
int ddleq_verify_dev(pgpu_ctx* ctx, size_t count, uint32_t secpar, const uint32_t* ct1, const uint32_t* ct2, const uint32_t* x, const uint32_t* y,
                     const uint32_t* alpha, const uint32_t* e, const uint32_t* f, uint8_t* ok) {
    if (!ctx->level2_ready) return fail(ctx, PGPU_ERR_UNSUPPORTED, "level 2: n^3 is wider than the built kernel shapes");
    if (secpar == 0) return fail(ctx, PGPU_ERR_ARG, "VerifyDDLEQProof: secpar must be positive");
    const ModCtx &M2 = ctx->m_n2, &M3 = ctx->m_n3;
    const uint32_t S2 = M2.sh.S, S3 = M3.sh.S, e2bits = (uint32_t)ctx->n2.bitlen();
    const size_t total = count * secpar;
    int rc;
    DEVBUF(dig, ctx, total * 8); DEVBUF(chk, ctx, total * S3); DEVBUF(en, ctx, total * S2); DEVBUF(fn2, ctx, total * S3); DEVBUF(t, ctx, total * S3);
    if ((rc = ddleq_hash(ctx, total, secpar, ct2, x, y, alpha, dig.p))) return rc;                                           // :138
    CU(ctx, select_launch(dig.p, ct2, secpar, ct1, secpar, S3, (uint32_t)total, chk.p, ctx->stream));                        // :140-143
    ctx->launches++;
    if ((rc = modexp_shared_dev(ctx, M2, total, e, ctx->n, en.p))) return rc;                                                // E^n        :145
    if (getenv("PGPU_NO_DUAL_EXP")) {
        if ((rc = modexp_shared_dev(ctx, M3, total, f, ctx->n2, fn2.p))) return rc;                                              // F^(n^2)    :146
        if ((rc = modexp_items_io(ctx, M3, total, IoDesc{chk.p, S3, S3}, ExpDesc{en.p, S2, e2bits, nullptr}, t.p))) return rc;  // check^en   :148
        if ((rc = modmul_dev(ctx, M3, total, t.p, fn2.p, t.p))) return rc;                                                       // :149-150
    } else {                                                                            // check^en * F^(n^2) with shared squarings :146-150
        if ((rc = modexp_dual_io(ctx, M3, total, IoDesc{chk.p, S3, S3}, ExpDesc{en.p, S2, e2bits, nullptr}, IoDesc{f, S3, S3}, ctx->n2, "n2", t.p))) return rc;
    }
    CU(ctx, equal_launch(alpha, t.p, S3, (uint32_t)total, ok, ctx->stream));                                                 // :152
    ctx->launches++;
    return PGPU_OK;
}

}  // namespace pgpu
