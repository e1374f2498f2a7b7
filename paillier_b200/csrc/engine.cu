// Engine internals (engine.hpp): key constants, program compiler, device-side operations.
#include "engine.hpp"

namespace pgpu {

namespace { thread_local std::string g_err; }
std::string& thread_error() { return g_err; }

// integer-pipe shape for records of S limbs (also what the 32-bit-limb helper kernels run on)
static bool int_shape(size_t limbs, Shape& out) {
    static const Shape defaults[] = {{32, 4, 8}, {64, 4, 16}, {96, 4, 24}, {128, 4, 32}, {192, 8, 24}};
    for (const Shape& s : defaults)
        if ((size_t)s.S >= limbs) { out = s; return true; }
    return false;
}

// Smallest built shape holding `limbs` limbs.  4096-bit moduli (n^2 of a 2048-bit key: EncryptWithR, PartialDecrypt) run on
// the FP64 pipe (52-bit limbs, mont52.cuh): +3 % over the integer-pipe kernel end to end.  The other widths stay on the
// integer pipe, where the FP64 kernel measured equal (2048-bit moduli, which have the dedicated squaring) or slower (3072-
// and 6144-bit moduli: 15 limbs per lane need 254 registers) -- profiles/r02_fp64_experiments.md.  PGPU_NO_FP64=1 keeps
// everything on the integer pipe; override per width with PGPU_SHAPE_<S>="tpi,L" (integer pipe) or "tpi,L,fp64".
bool pick_shape(size_t limbs, Shape& out) {
    static const Shape fp64_defaults[] = {{128, 8, 10, true}};
    static const bool no_fp64 = getenv("PGPU_NO_FP64") != nullptr;
    if (!int_shape(limbs, out)) return false;
    if (!no_fp64)
        for (const Shape& s : fp64_defaults)
            if (s.S == out.S && vm_occupancy(s) > 0) { out = s; break; }
    char name[32];
    snprintf(name, sizeof name, "PGPU_SHAPE_%d", out.S);
    if (const char* e = getenv(name)) {
        int t = 0, l = 0; char tag[8] = "";
        const int got = sscanf(e, "%d,%d,%7s", &t, &l, tag);
        Shape c{out.S, t, l, got == 3 && std::string(tag) == "fp64"};
        if (got >= 2 && (c.fp64 || t * l == out.S) && vm_occupancy(c) > 0) out = c;
    }
    return true;
}

int fail(pgpu_ctx* ctx, int code, const std::string& msg) {
    g_err = msg;
    if (ctx) ctx->err = msg;
    return code;
}

int upload(pgpu_ctx* ctx, uint32_t* dst, const std::vector<uint32_t>& v) {
    CU(ctx, cudaMemcpyAsync(dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return PGPU_OK;
}

int set_kconst(pgpu_ctx* ctx, ModCtx& m, uint32_t slot, const BigU& v) {
    return upload(ctx, m.d_kconst + (size_t)slot * m.sh.S, v.limbs(m.sh.S));
}

int modctx_init(pgpu_ctx* ctx, ModCtx& m, const BigU& N) {
    if (!N.is_odd()) return fail(ctx, PGPU_ERR_ARG, "modulus must be odd");
    if (!pick_shape(N.v.size(), m.sh)) return fail(ctx, PGPU_ERR_UNSUPPORTED, "modulus wider than the built kernel shapes");
    m.N = N;
    int_shape(N.v.size(), m.sh32);
    const BigU R = BigU::pow2((size_t)m.sh.rbits());
    m.R1 = R % N;
    m.R2 = (m.R1 * m.R1) % N;
    m.W1 = BigU::pow2(32 * (size_t)m.sh.S) % N;
    m.R3 = (m.R2 * m.W1) % N;              // takes the high chunk of a record split at 2^(32*S) into Montgomery form
    m.np0 = mont_np0(N.v[0]);
    CU(ctx, cudaMalloc(&m.d_mod, (size_t)m.sh.S * 4));
    CU(ctx, cudaMalloc(&m.d_kconst, (size_t)K_SLOTS * m.sh.S * 4));
    CU(ctx, cudaMemsetAsync(m.d_kconst, 0, (size_t)K_SLOTS * m.sh.S * 4, ctx->stream));
    int rc;
    if ((rc = upload(ctx, m.d_mod, N.limbs(m.sh.S)))) return rc;
    if ((rc = set_kconst(ctx, m, K_R2, m.R2))) return rc;
    if ((rc = set_kconst(ctx, m, K_R1, m.R1))) return rc;
    if ((rc = set_kconst(ctx, m, K_ONE, BigU(1)))) return rc;
    if ((rc = set_kconst(ctx, m, K_R3, m.R3))) return rc;
    m.blocks_per_sm = vm_occupancy(m.sh);
    if (m.blocks_per_sm <= 0) return fail(ctx, PGPU_ERR_CUDA, "powm_vm occupancy query failed for shape");
    m.sh_items = m.sh; m.blocks_per_sm_items = m.blocks_per_sm;
    if (!m.sh.fp64 && m.sh.S == 96 && m.sh.tpi == 4 && !getenv("PGPU_SHAPE_96")) {      // measured: tools/ddleq_rate.py, tools/shape96.py
        const int b = vm_occupancy(Shape{96, 8, 12});
        if (b > 0) { m.sh_items = Shape{96, 8, 12}; m.blocks_per_sm_items = b; }
    }
    m.ready = true;
    return PGPU_OK;
}

void dev_scrub_free(void* p, size_t bytes) {
    if (!p) return;
    if (bytes) {
        cudaDeviceSynchronize();
        cudaMemset(p, 0, bytes);
    }
    cudaFree(p);
}

void scrub(std::vector<uint32_t>& v) {
    volatile uint32_t* q = v.data();
    for (size_t i = 0; i < v.size(); ++i) q[i] = 0;
    v.clear();
}

void scrub(BigU& x) { scrub(x.v); }

void modctx_free(ModCtx& m) {
    dev_scrub_free(m.d_mod, (size_t)m.sh.S * 4);
    dev_scrub_free(m.d_kconst, (size_t)K_SLOTS * m.sh.S * 4);
    scrub(m.N); scrub(m.R1); scrub(m.R2); scrub(m.R3); scrub(m.W1);
    m = ModCtx();
}

// ------------------------------------------------------- program compiler
int choose_window(size_t bits) {
    int best = 1; double best_cost = 1e300;
    static const int wmax = [] { const char* e = getenv("PGPU_WINDOW_MAX"); const int v = e ? atoi(e) : 7; return v < 1 ? 1 : v > 7 ? 7 : v; }();
    for (int w = 1; w <= wmax; ++w) {
        double cost = (w == 1 ? 0.0 : (double)(1u << (w - 1))) + (double)bits / (w + 1);
        if (cost < best_cost) { best_cost = cost; best = w; }
    }
    return best;
}

// x holds the base in Montgomery form; afterwards x = base^e (Montgomery form).
// Sliding window over the shared exponent e: odd powers in T[tb .. tb+2^(w-1)),
// base^2 in T[tb+2^(w-1)].   gmp semantics: e = 0 gives 1.
void emit_pow_shared(Program& P, const BigU& e, uint32_t tb) {
    const size_t bits = e.bitlen();
    if (bits == 0) { P.emit(OP_LDC, K_R1); return; }
    const int w = choose_window(bits);
    const uint32_t nodd = 1u << (w - 1);
    P.emit(OP_STT, tb); P.use_slot(tb);
    if (w > 1) {
        const uint32_t sq = tb + nodd;
        P.emit(OP_SQR, 1); P.n_sqr += 1;
        P.emit(OP_STT, sq); P.use_slot(sq);
        for (uint32_t i = 1; i < nodd; ++i) {   // x = T[i-1] * base^2
            if (i == 1) P.emit(OP_LDT, tb);
            P.emit(OP_MULT, sq); P.n_mul += 1;
            P.emit(OP_STT, tb + i);
        }
    }
    bool first = true;
    uint32_t pending = 0;
    long i = (long)bits - 1;
    while (i >= 0) {
        if (!e.bit(i)) { ++pending; --i; continue; }
        long j = std::max<long>(i - w + 1, 0);
        while (!e.bit(j)) ++j;
        uint32_t val = 0;
        for (long k = i; k >= j; --k) val = (val << 1) | (e.bit(k) ? 1u : 0u);
        const uint32_t len = (uint32_t)(i - j + 1), idx = tb + (val >> 1);
        if (first) {
            P.emit(OP_LDT, idx);
            first = false;
        } else {
            uint32_t nsq = pending + len;
            if (nsq > 0xfff) { P.emit(OP_SQR, nsq - len); nsq = len; }
            P.emit(OP_SQMT, nsq | (idx << 12));
            P.n_sqr += pending + len; P.n_mul += 1;
        }
        pending = 0;
        i = j - 1;
    }
    if (pending) { P.emit(OP_SQR, pending); P.n_sqr += pending; }
}

// per-item exponents of exp_bits bits (OP_WIN): full table T[tb .. tb+2^w), fixed window
void emit_pow_items(Program& P, size_t exp_bits, uint32_t tb) {
    int w = 1;
    { double best = 1e300; for (int c = 1; c <= 6; ++c) { double cost = (double)(1u << c) + (double)exp_bits / c; if (cost < best) { best = cost; w = c; } } }
    const uint32_t n = 1u << w;
    // T[tb+1] = base, T[tb] = 1, T[tb+i] = T[tb+i-1] * base
    P.emit(OP_STT, tb + 1);
    for (uint32_t i = 2; i < n; ++i) { P.emit(OP_MULT, tb + 1); P.n_mul += 1; P.emit(OP_STT, tb + i); }
    P.emit(OP_LDC, K_R1); P.emit(OP_STT, tb);
    P.use_slot(tb + n - 1);
    const size_t nwin = (exp_bits + w - 1) / w;
    for (size_t k = nwin; k-- > 0;) {
        // x starts at 1, so the first squarings are of 1; keep the program uniform
        P.emit(OP_WIN, (uint32_t)(k * w) | ((uint32_t)w << 20) | (tb << 24));
        P.n_sqr += w; P.n_mul += 1;
    }
}

int program_upload(pgpu_ctx* ctx, Program& P) {
    P.ops.push_back(vm_op(OP_END, 0));
    dev_scrub_free(P.d_ops, P.d_bytes);
    P.d_ops = nullptr; P.d_bytes = 0;
    CU(ctx, cudaMalloc(&P.d_ops, P.ops.size() * 4));
    P.d_bytes = P.ops.size() * 4;
    return upload(ctx, P.d_ops, P.ops);
}

void program_free(Program& P) { dev_scrub_free(P.d_ops, P.d_bytes); scrub(P.ops); P = Program(); }

// ------------------------------------------------------------- launching
int ensure_table(pgpu_ctx* ctx, size_t limbs) {
    if (limbs <= ctx->table_limbs) return PGPU_OK;
    if (ctx->d_table) { CU(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_table); ctx->d_table = nullptr; ctx->table_limbs = 0; }
    CU(ctx, cudaMalloc(&ctx->d_table, limbs * 4));
    ctx->table_limbs = limbs;
    return PGPU_OK;
}

int run_vm(pgpu_ctx* ctx, const ModCtx& m, const Program& prog, size_t count,
           const IoDesc* ins, int n_in, uint32_t* out, uint32_t out_stride, uint32_t out_limbs,
           const ExpDesc& ex, uint32_t* out2, uint32_t out2_stride, int force_blocks) {
    if (count == 0) return PGPU_OK;
    if (count > 0x7fffffffu) return fail(ctx, PGPU_ERR_ARG, "batch too large");
    const bool items_shape = prog.per_item && force_blocks <= 0;
    const Shape& sh = items_shape ? m.sh_items : m.sh;
    const int gpb = VM_BLOCK_THREADS / sh.tpi;
    const size_t max_blocks = (size_t)ctx->sms * (items_shape ? m.blocks_per_sm_items : m.blocks_per_sm);
    const size_t want = (count + gpb - 1) / gpb;
    const int blocks = force_blocks > 0 ? force_blocks : (int)std::min(max_blocks, want);
    VmParams P{};
    P.prog = prog.d_ops;
    P.n_items = (uint32_t)count;
    P.mod = m.d_mod; P.np0 = m.np0; P.kconst = m.d_kconst;
    for (int i = 0; i < VM_MAX_IN; ++i) P.in_div[i] = 1;
    for (int i = 0; i < n_in; ++i) { P.in[i] = ins[i].ptr; P.in_stride[i] = ins[i].stride; P.in_limbs[i] = ins[i].limbs; P.in_div[i] = std::max<uint32_t>(ins[i].div, 1); }
    P.out[0] = out; P.out_stride[0] = out_stride; P.out_limbs[0] = out_limbs;
    P.out[1] = out2; P.out_stride[1] = out2_stride; P.out_limbs[1] = m.sh.S;
    P.exp = ex.ptr; P.exp_stride = ex.stride; P.exp_bits = ex.bits; P.fixed = ex.fixed; P.exp_sub = ex.sub;
    P.n_groups = (uint32_t)blocks * gpb;
    { static const bool no_sqr = getenv("PGPU_NO_SQR") != nullptr; P.flags = no_sqr ? 1u : 0u; }
    const size_t tbl_limbs = (size_t)std::max<uint32_t>(prog.tbl_entries, 1) * P.n_groups * sh.tbl_limbs();
    int rc = ensure_table(ctx, tbl_limbs + (size_t)P.n_groups * m.sh.S);
    if (rc) return rc;
    P.table = ctx->d_table;
    P.dump = ctx->d_table + tbl_limbs;
    CU(ctx, vm_launch(sh, P, blocks, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

int vm_full_blocks(const pgpu_ctx* ctx, const ModCtx& m) { return ctx->sms * m.blocks_per_sm; }

int stage(pgpu_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->stage_bytes[slot] < bytes) {
        if (ctx->d_stage[slot]) { CU(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_stage[slot]); ctx->d_stage[slot] = nullptr; ctx->stage_bytes[slot] = 0; }
        CU(ctx, cudaMalloc(&ctx->d_stage[slot], std::max<size_t>(bytes, 256)));
        ctx->stage_bytes[slot] = std::max<size_t>(bytes, 256);
    }
    *out = ctx->d_stage[slot];
    return PGPU_OK;
}

ModCtx* select_mod(pgpu_ctx* ctx, int modsel) {
    switch (modsel) {
        case PGPU_MOD_N: return ctx->m_n.ready ? &ctx->m_n : nullptr;
        case PGPU_MOD_N2: return ctx->m_n2.ready ? &ctx->m_n2 : nullptr;
        case PGPU_MOD_N3: return ctx->m_n3.ready ? &ctx->m_n3 : nullptr;
    }
    return nullptr;
}

// ---------------------------------------------------- compiled key programs
// EncryptWithR, level 1 (paillier.go:206-218) with the g = n+1 shortcut:
// c = (1 + m*n) * r^n mod n^2.   in0 = r, in1 = m.
int build_encrypt(pgpu_ctx* ctx) {
    Program& P = ctx->prog_enc;
    program_free(P);
    P.emit(OP_LDI, 0);
    P.emit(OP_MULC, K_R2); P.n_mul++;           // r -> Montgomery form
    emit_pow_shared(P, ctx->n, 0);              // r^n
    const uint32_t RES = P.tbl_entries;         // next free table slot
    P.emit(OP_STT, RES); P.use_slot(RES);
    P.emit(OP_LDI, 1);
    P.emit(OP_MULC, K_NR2); P.n_mul++;          // m*n (Montgomery form), exact since m*n < n^2
    P.emit(OP_ADDC, K_R1);                      // + 1
    P.emit(OP_MULT, RES); P.n_mul++;
    P.emit(OP_MULC, K_ONE); P.n_mul++;          // out of Montgomery form
    P.emit(OP_STO, 0);
    return program_upload(ctx, P);
}

// One CRT half of Decrypt: x = c^(p-1) mod p^2.  in0 = low half of c, in1 = high half
// (c = lo + hi*R with R = 2^(32*S) of the p^2 shape).
int build_decrypt_half(pgpu_ctx* ctx, Program& P, const BigU& pm1) {
    program_free(P);
    P.emit(OP_LDI, 0);
    P.emit(OP_MULC, K_R2); P.n_mul++;           // lo*R
    P.emit(OP_STT, 0); P.use_slot(0);
    P.emit(OP_LDI, 1);
    P.emit(OP_MULC, K_R3); P.n_mul++;           // hi*R*R
    P.emit(OP_ADDT, 0);                         // (c mod p^2) in Montgomery form
    emit_pow_shared(P, pm1, 0);
    P.emit(OP_MULC, K_ONE); P.n_mul++;
    P.emit(OP_STO, 0);
    return program_upload(ctx, P);
}

// PartialDecrypt (thresholdkey.go:192-201): c^(2*delta*share) mod n^2
int build_pdec(pgpu_ctx* ctx) {
    Program& P = ctx->prog_pdec;
    program_free(P);
    const BigU e = ctx->tk_share * (BigU(2) * ctx->tk_delta);
    P.emit(OP_LDI, 0);
    P.emit(OP_MULC, K_R2); P.n_mul++;
    emit_pow_shared(P, e, 0);
    P.emit(OP_MULC, K_ONE); P.n_mul++;
    P.emit(OP_STO, 0);
    return program_upload(ctx, P);
}

// EncryptWithR for the holder of p, q (paillier.go:206-218 gives the same c for every caller; SecretKey embeds
// PublicKey, paillier.go:29-34): r^n is computed mod q^2 and mod p^2 (two half-width exponentiations, 1/4 of the
// n^2 cost each), recombined with Garner's formula x = x_q + q^2 * ((x_p - x_q) * q^-2 mod p^2), and multiplied
// by g^m = 1 + m*n mod n^2.  No coprimality assumption: the exponent n is used unreduced.
int setup_encrypt_crt(pgpu_ctx* ctx) {
    ctx->has_enc_crt = false;
    ModCtx &P2 = ctx->m_p2, &N2 = ctx->m_n2;
    if (ctx->wn > (size_t)P2.sh.S || (size_t)P2.sh.S > (size_t)N2.sh.S) return PGPU_OK;   // r does not fit a p^2 record: general path
    const BigU p2 = ctx->p * ctx->p, q2 = ctx->q * ctx->q;
    BigU q2inv;
    if (!BigU::modinv(q2 % p2, p2, q2inv)) return fail(ctx, PGPU_ERR_ARG, "p and q are not coprime");
    int rc;
    if ((rc = set_kconst(ctx, P2, K_CRT, q2inv))) return rc;
    if ((rc = set_kconst(ctx, N2, K_CRT, (q2 * N2.R2) % N2.N))) return rc;
    {   // x_q = r^n mod q^2.  in0 = r
        Program& P = ctx->prog_encq;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_R2); P.n_mul++;       // r < n may exceed q^2: the Montgomery product reduces it
        emit_pow_shared(P, ctx->n, 0);
        P.emit(OP_MULC, K_ONE); P.n_mul++;
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    {   // t = (r^n - x_q) * q^-2 mod p^2.  in0 = r, in1 = x_q
        Program& P = ctx->prog_encp;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        emit_pow_shared(P, ctx->n, 0);
        const uint32_t XP = P.tbl_entries, XQ = XP + 1;
        P.emit(OP_STT, XP); P.use_slot(XP);
        P.emit(OP_LDI, 1);
        P.emit(OP_MULC, K_R2); P.n_mul++;       // x_q mod p^2, Montgomery form
        P.emit(OP_STT, XQ); P.use_slot(XQ);
        P.emit(OP_LDT, XP);
        P.emit(OP_SUBT, XQ);
        P.emit(OP_MULC, K_CRT); P.n_mul++;      // times the plain constant: leaves Montgomery form
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    {   // c = (x_q + q^2*t) * (1 + m*n) mod n^2.  in0 = t, in1 = x_q, in2 = m
        Program& P = ctx->prog_encf;
        program_free(P);
        P.emit(OP_LDI, 0);
        P.emit(OP_MULC, K_CRT); P.n_mul++;      // q^2*t, Montgomery form
        P.emit(OP_STT, 0); P.use_slot(0);
        P.emit(OP_LDI, 1);
        P.emit(OP_MULC, K_R2); P.n_mul++;
        P.emit(OP_ADDT, 0);                     // r^n mod n^2 (x_q + q^2*t < n^2: the modular add is exact)
        P.emit(OP_STT, 0);
        P.emit(OP_LDI, 2);
        P.emit(OP_MULC, K_NR2); P.n_mul++;
        P.emit(OP_ADDC, K_R1);                  // g^m = 1 + m*n
        P.emit(OP_MULT, 0); P.n_mul++;
        P.emit(OP_MULC, K_ONE); P.n_mul++;
        P.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, P))) return rc;
    }
    ctx->has_enc_crt = true;
    return PGPU_OK;
}

int setup_crt(pgpu_ctx* ctx) {
    const BigU &p = ctx->p, &q = ctx->q;
    int rc;
    modctx_free(ctx->m_p2); modctx_free(ctx->m_q2);
    if ((rc = modctx_init(ctx, ctx->m_p2, p * p))) return rc;
    if ((rc = modctx_init(ctx, ctx->m_q2, q * q))) return rc;
    if (ctx->m_p2.sh.S != ctx->m_q2.sh.S) return fail(ctx, PGPU_ERR_UNSUPPORTED, "p and q of different width");
    if ((rc = build_decrypt_half(ctx, ctx->prog_dec_p, p - BigU(1)))) return rc;
    if ((rc = build_decrypt_half(ctx, ctx->prog_dec_q, q - BigU(1)))) return rc;
    const int h = ctx->m_p2.sh.S / 2;
    if (h > CRT_MAXH || p.v.size() > (size_t)h || q.v.size() > (size_t)h) return fail(ctx, PGPU_ERR_UNSUPPORTED, "prime factors too wide for the CRT recombination");
    ctx->crt_h = h;
    const BigU Rh = BigU::pow2(32 * (size_t)h);
    BigU pinv, qinv, qinv_p, hp, hq, t;
    if (!BigU::modinv(p, Rh, pinv) || !BigU::modinv(q, Rh, qinv)) return fail(ctx, PGPU_ERR_ARG, "p, q must be odd");
    if (!BigU::modinv(q % p, p, qinv_p)) return fail(ctx, PGPU_ERR_ARG, "p and q are not coprime");
    // h_p = L_p(g^(p-1) mod p^2)^-1 mod p with g = n+1: g^(p-1) = 1 + (p-1)*n mod p^2
    const BigU p2 = p * p, q2 = q * q;
    BigU gp = (BigU(1) + (p - BigU(1)) * ctx->n) % p2, gq = (BigU(1) + (q - BigU(1)) * ctx->n) % q2;
    BigU lp = ((gp - BigU(1)) / p) % p, lq = ((gq - BigU(1)) / q) % q;
    if (!BigU::modinv(lp, p, hp) || !BigU::modinv(lq, q, hq)) return fail(ctx, PGPU_ERR_ARG, "invalid key: L(g^(p-1)) not invertible");
    std::vector<uint32_t> K;
    auto push = [&](const BigU& x) { auto l = x.limbs(h); K.insert(K.end(), l.begin(), l.end()); };
    push(p); push(q); push(pinv); push(qinv);
    push((hp * Rh) % p); push((hq * Rh) % q); push((qinv_p * Rh) % p);
    dev_scrub_free(ctx->d_crt, ctx->d_crt_bytes);
    ctx->d_crt = nullptr; ctx->d_crt_bytes = 0;
    CU(ctx, cudaMalloc(&ctx->d_crt, K.size() * 4));
    ctx->d_crt_bytes = K.size() * 4;
    rc = upload(ctx, ctx->d_crt, K);
    scrub(K);
    if (rc) return rc;
    ctx->crt_np0_p = mont_np0(p.v[0]); ctx->crt_np0_q = mont_np0(q.v[0]);
    ctx->has_secret = true;
    if ((rc = setup_encrypt_crt(ctx))) return rc;
    return setup_level2_secret(ctx);
}

// ------------------------------------------------------- device-side ops
int encrypt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c) {
    const ModCtx& M = ctx->m_n2;
    IoDesc ins[2] = {{r, (uint32_t)ctx->wn, (uint32_t)ctx->wn}, {m, (uint32_t)ctx->wn, (uint32_t)ctx->wn}};
    return run_vm(ctx, M, ctx->prog_enc, count, ins, 2, c, M.sh.S, M.sh.S);
}

int decrypt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, uint32_t* m) {
    if (!ctx->has_secret) return fail(ctx, PGPU_ERR_STATE, "Decrypt: no secret key loaded");
    const ModCtx &P2 = ctx->m_p2, &Q2 = ctx->m_q2;
    const uint32_t S2 = ctx->m_n2.sh.S, Sp = P2.sh.S;
    // c (S2 limbs) = lo (Sp limbs) + hi * 2^(32*Sp); S2 <= 2*Sp
    const uint32_t hi_limbs = S2 > Sp ? S2 - Sp : 0;
    void *xp, *xq; int rc;
    if ((rc = stage(ctx, 12, count * Sp * 4, &xp))) return rc;
    if ((rc = stage(ctx, 13, count * Sp * 4, &xq))) return rc;
    IoDesc ins[2] = {{c, S2, std::min(S2, Sp)}, {c + Sp, S2, hi_limbs}};
    if ((rc = run_vm(ctx, P2, ctx->prog_dec_p, count, ins, 2, (uint32_t*)xp, Sp, Sp))) return rc;
    if ((rc = run_vm(ctx, Q2, ctx->prog_dec_q, count, ins, 2, (uint32_t*)xq, Sp, Sp))) return rc;
    CrtParams C{};
    C.n_items = (uint32_t)count; C.h = ctx->crt_h; C.consts = ctx->d_crt;
    C.np0_p = ctx->crt_np0_p; C.np0_q = ctx->crt_np0_q;
    C.xp = (const uint32_t*)xp; C.xq = (const uint32_t*)xq; C.x_stride = Sp;
    C.out = m; C.out_stride = (uint32_t)ctx->wn; C.out_limbs = (uint32_t)ctx->wn;
    CU(ctx, crt_combine_launch(C, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

// Online half of an offline/online encryption: c = (1 + m*n) * rn mod n^2 with rn = r^n mod n^2 computed ahead of
// time (EncryptWithR with m = 0 yields exactly r^n).  Same c as EncryptWithR(m, r), paillier.go:206-218.
int encrypt_rn_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* rn, uint32_t* c) {
    const ModCtx& M = ctx->m_n2;
    const std::string key = "encrn";
    Program* P = cached_program(ctx, key);
    if (!P) {
        Program np;
        np.emit(OP_LDI, 0);
        np.emit(OP_MULC, K_NR2); np.n_mul++;       // m*n, Montgomery form
        np.emit(OP_ADDC, K_R1);                    // + 1
        np.emit(OP_MULI, 1); np.n_mul++;           // times the plain rn: leaves Montgomery form
        np.emit(OP_STO, 0);
        int rc = program_upload(ctx, np);
        if (rc) return rc;
        P = &(ctx->prog_cache[key] = np);
    }
    IoDesc ins[2] = {{m, (uint32_t)ctx->wn, (uint32_t)ctx->wn}, {rn, (uint32_t)M.sh.S, (uint32_t)M.sh.S}};
    return run_vm(ctx, M, *P, count, ins, 2, c, M.sh.S, M.sh.S);
}

int encrypt_crt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c) {
    if (!ctx->has_secret) return fail(ctx, PGPU_ERR_STATE, "EncryptWithR (secret key): no secret key loaded");
    if (!ctx->has_enc_crt) return encrypt_dev(ctx, count, m, r, c);
    const ModCtx &P2 = ctx->m_p2, &Q2 = ctx->m_q2, &N2 = ctx->m_n2;
    const uint32_t Sp = P2.sh.S, wn = (uint32_t)ctx->wn;
    void *xq, *t; int rc;
    if ((rc = stage(ctx, 12, count * Sp * 4, &xq))) return rc;
    if ((rc = stage(ctx, 13, count * Sp * 4, &t))) return rc;
    IoDesc iq[1] = {{r, wn, wn}};
    if ((rc = run_vm(ctx, Q2, ctx->prog_encq, count, iq, 1, (uint32_t*)xq, Sp, Sp))) return rc;
    IoDesc ip[2] = {{r, wn, wn}, {(const uint32_t*)xq, Sp, Sp}};
    if ((rc = run_vm(ctx, P2, ctx->prog_encp, count, ip, 2, (uint32_t*)t, Sp, Sp))) return rc;
    IoDesc fin[3] = {{(const uint32_t*)t, Sp, Sp}, {(const uint32_t*)xq, Sp, Sp}, {m, wn, wn}};
    return run_vm(ctx, N2, ctx->prog_encf, count, fin, 3, c, N2.sh.S, N2.sh.S);
}

int pdec_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, uint32_t* out) {
    if (!ctx->has_share) return fail(ctx, PGPU_ERR_STATE, "PartialDecrypt: no threshold share loaded");
    const ModCtx& M = ctx->m_n2;
    IoDesc ins[1] = {{c, (uint32_t)M.sh.S, (uint32_t)M.sh.S}};
    return run_vm(ctx, M, ctx->prog_pdec, count, ins, 1, out, M.sh.S, M.sh.S);
}

Program* cached_program(pgpu_ctx* ctx, const std::string& key) {
    auto it = ctx->prog_cache.find(key);
    return it == ctx->prog_cache.end() ? nullptr : &it->second;
}

// out[i] = base[i]^exp[i] mod M with per-item exponents of exp.bits bits (fixed window)
int modexp_items_io(pgpu_ctx* ctx, const ModCtx& M, size_t count, const IoDesc& base, const ExpDesc& exp, uint32_t* out) {
    const std::string key = "powi:" + std::to_string(M.sh.S) + ":" + std::to_string(exp.bits);
    Program* P = cached_program(ctx, key);
    if (!P) {
        Program np;
        np.emit(OP_LDI, 0);
        np.emit(OP_MULC, K_R2); np.n_mul++;
        emit_pow_items(np, exp.bits, 0);
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STO, 0);
        int rc = program_upload(ctx, np);
        if (rc) return rc;
        P = &(ctx->prog_cache[key] = np);
    }
    IoDesc ins[1] = {base};
    return run_vm(ctx, M, *P, count, ins, 1, out, M.sh.S, M.sh.S, exp);
}

int modexp_items_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* base, const uint32_t* exp, uint32_t exp_limbs, uint32_t* out,
                     bool broadcast_base) {
    return modexp_items_io(ctx, M, count, IoDesc{base, broadcast_base ? 0u : (uint32_t)M.sh.S, (uint32_t)M.sh.S},
                           ExpDesc{exp, exp_limbs, 32 * exp_limbs, nullptr}, out);
}

// Shared-base multi-exponentiation: k exponents per base, item-major (record i*k + s).  VerifyProof of the k share-holders'
// proofs for one ciphertext raises the same c^4 to k different Z (thresholdkey.go:293-302): right to left, the chain
// c^4, (c^4)^(2^w), (c^4)^(2^(2w)), ... is squared ONCE per ciphertext and every exponent s multiplies it into the bucket of its
// digit (one multiplication per window and exponent, OP_BKT with sub = s); per exponent the buckets are then folded into
// prod_d T[d]^d with the running-product trick.  nwin*w squarings + k*(nwin + 2*2^w) multiplications per base instead of
// k*(nwin*w + nwin + 2^w): 3.7x fewer for k = 8 at 6656-bit exponents.
// window width of the bucket method for k exponents of `bits` bits
// (w <= 6: 8 x 65 buckets of 768 B per resident group are 1.9 GB at 6144 bits; w = 7 would save 1 % for twice that)
static int multi_window(uint32_t k, uint32_t bits) {
    int w = 1;
    double best = 1e300;
    for (int c = 1; c <= 6; ++c) { const double cost = (double)k * ((double)bits / c + 2.0 * (1u << c)); if (cost < best) { best = cost; w = c; } }
    return w;
}

// in0 = base (plain); out[0][item*k + s] = (base^pre)^(exp_s), exp_s = exponent s of the item (OP_BKT with sub = s)
void build_multi_program(Program& np, uint32_t k, uint32_t bits, uint32_t pre) {
    const int w = multi_window(k, bits);
    const uint32_t nb = 1u << w, per = nb + 1, nwin = (bits + w - 1) / w;
    const uint32_t CH = k * per, RUN = CH + 1, ACC = CH + 2;
    np.emit(OP_LDC, K_R1);
    for (uint32_t s = 0; s < k; ++s) for (uint32_t d = 0; d <= nb; ++d) np.emit(OP_STT, s * per + d);       // every bucket = 1
    np.use_slot(ACC);
    np.emit(OP_LDI, 0);
    np.emit(OP_MULC, K_R2); np.n_mul++;
    if (pre == 2) { np.emit(OP_SQR, 1); np.n_sqr += 1; }
    if (pre == 4) { np.emit(OP_SQR, 2); np.n_sqr += 2; }
    for (uint32_t win = 0; win < nwin; ++win) {
        np.emit(OP_STT, CH);                                                              // the chain value of this window
        for (uint32_t s = 0; s < k; ++s) {
            np.emit(OP_BKT, (win * w) | ((uint32_t)w << 20) | (s << 24)); np.n_mul++;     // T[s][digit] *= chain
            if (s + 1 < k || win + 1 < nwin) np.emit(OP_LDT, CH);
        }
        if (win + 1 < nwin) { np.emit(OP_SQR, (uint32_t)w); np.n_sqr += w; }
    }
    for (uint32_t s = 0; s < k; ++s) {
        // run = T[nb-1]; acc = run; for d = nb-2 .. 1: run *= T[d]; acc *= run          (acc = prod_d T[d]^d)
        np.emit(OP_LDT, s * per + nb - 1);
        if (nb > 2) {
            np.emit(OP_STT, RUN); np.emit(OP_STT, ACC);
            for (uint32_t d = nb - 2; d >= 1; --d) {
                np.emit(OP_LDT, RUN); np.emit(OP_MULT, s * per + d); np.n_mul++; np.emit(OP_STT, RUN);
                np.emit(OP_MULT, ACC); np.n_mul++; np.emit(OP_STT, ACC);
            }
        }
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STOO, (s << 2) | 0u);
    }
}

int modexp_multi_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, uint32_t k, const uint32_t* base, uint32_t pre, const uint32_t* exp, uint32_t exp_limbs,
                     uint32_t* out) {
    if (k == 0 || count == 0) return PGPU_OK;
    if (k > 8) return fail(ctx, PGPU_ERR_ARG, "modexp_multi: at most 8 exponents per base");
    const uint32_t S = M.sh.S, bits = 32 * exp_limbs;
    const std::string key = "multi:" + std::to_string(S) + ":" + std::to_string(k) + ":" + std::to_string(bits) + ":" + std::to_string(pre);
    Program* P = cached_program(ctx, key);
    if (!P) {
        Program np;
        build_multi_program(np, k, bits, pre);
        int rc = program_upload(ctx, np);
        if (rc) return rc;
        P = &(ctx->prog_cache[key] = np);
    }
    IoDesc ins[1] = {{base, S, S}};
    ExpDesc ex{exp, k * exp_limbs, bits, nullptr, exp_limbs};
    return run_vm(ctx, M, *P, count, ins, 1, out, k * S, S, ex);
}

// out[i] = base[i]^e mod M, one exponent for the whole batch (sliding window compiled on the host)
int modexp_shared_io(pgpu_ctx* ctx, const ModCtx& M, size_t count, const IoDesc& base, const BigU& e, uint32_t* out) {
    // Ad-hoc exponents (a combine compiles one program per share and subset: 2*lambda_i) live in a small LRU of their own, so
    // a workload with many distinct exponents never evicts the long-lived programs (mul, fix, powi, xr-*, ...) of prog_cache.
    constexpr size_t POWS_CAPACITY = 256;
    const std::string key = "pows:" + std::to_string(M.sh.S) + ":" + e.hex();
    Program* P = cached_program(ctx, key);
    if (P) {
        ctx->pows_lru.remove(key);
        ctx->pows_lru.push_back(key);
    } else {
        Program np;
        np.emit(OP_LDI, 0);
        np.emit(OP_MULC, K_R2); np.n_mul++;
        emit_pow_shared(np, e, 0);
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STO, 0);
        int rc = program_upload(ctx, np);
        if (rc) return rc;
        while (ctx->pows_lru.size() >= POWS_CAPACITY) {
            const std::string victim = ctx->pows_lru.front();
            ctx->pows_lru.pop_front();
            auto it = ctx->prog_cache.find(victim);
            if (it != ctx->prog_cache.end()) {
                CU(ctx, cudaStreamSynchronize(ctx->stream));        // a queued launch may still read the victim's ops
                if (it->second.d_ops) cudaFree(it->second.d_ops);
                ctx->prog_cache.erase(it);
            }
        }
        P = &(ctx->prog_cache[key] = np);
        ctx->pows_lru.push_back(key);
    }
    IoDesc ins[1] = {base};
    return run_vm(ctx, M, *P, count, ins, 1, out, M.sh.S, M.sh.S);
}

int modexp_shared_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* base, const BigU& e, uint32_t* out) {
    return modexp_shared_io(ctx, M, count, IoDesc{base, (uint32_t)M.sh.S, (uint32_t)M.sh.S}, e, out);
}

int modmul_io(pgpu_ctx* ctx, const ModCtx& M, size_t count, const IoDesc& a, const IoDesc& b, uint32_t* out) {
    const std::string key = "mul:" + std::to_string(M.sh.S);
    Program* P = cached_program(ctx, key);
    if (!P) {
        Program np;
        np.emit(OP_LDI, 0);
        np.emit(OP_MULC, K_R2); np.n_mul++;
        np.emit(OP_MULI, 1); np.n_mul++;
        np.emit(OP_STO, 0);
        int rc = program_upload(ctx, np);
        if (rc) return rc;
        P = &(ctx->prog_cache[key] = np);
    }
    IoDesc ins[2] = {a, b};
    return run_vm(ctx, M, *P, count, ins, 2, out, M.sh.S, M.sh.S);
}

int modmul_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    return modmul_io(ctx, M, count, IoDesc{a, (uint32_t)M.sh.S, (uint32_t)M.sh.S}, IoDesc{b, (uint32_t)M.sh.S, (uint32_t)M.sh.S}, out);
}

// out = prod in[i] mod M (Add over a batch)
int prod_dev(pgpu_ctx* ctx, ModCtx& M, size_t count, const uint32_t* in, uint32_t* out) {
    const int gpb = 128 / M.sh32.tpi;
    const size_t max_blocks = (size_t)ctx->sms * 4;
    size_t b1 = std::min(max_blocks, (count + gpb - 1) / gpb);
    if (b1 == 0) b1 = 1;
    void* part; int rc;
    if ((rc = stage(ctx, 14, (b1 + 1) * M.sh.S * 4, &part))) return rc;
    uint32_t* partial = (uint32_t*)part;
    auto rounds = [&](size_t n, size_t blocks) { size_t G = blocks * gpb; return n == 0 ? (size_t)1 : (n + G - 1) / G; };
    // stage 1: b1 blocks -> b1 partials, each carrying R^-(gpb*rounds1 - 1)
    ProdParams P1{in, (uint32_t)count, M.d_mod, M.np0, partial};
    CU(ctx, prod_reduce_launch(M.sh32.tpi, M.sh32.L, P1, (int)b1, ctx->stream));
    ctx->launches++;
    uint64_t T = (uint64_t)b1 * (gpb * rounds(count, b1) - 1);
    uint32_t* last = partial;
    if (b1 > 1) {
        uint32_t* fin = partial + b1 * M.sh.S;
        ProdParams P2{partial, (uint32_t)b1, M.d_mod, M.np0, fin};
        CU(ctx, prod_reduce_launch(M.sh32.tpi, M.sh32.L, P2, 1, ctx->stream));
        ctx->launches++;
        T += gpb * rounds(b1, 1) - 1;
        last = fin;
    }
    // final correction: the partial product carries W^-T (W = 2^(32*S), the radix of prod_reduce_kernel); one more
    // Montgomery multiply of the exponentiation kernel (radix R) by W^T * R restores it
    if (M.fix_T != T + 1) {                         // the constant depends on the batch geometry only: keep it across calls
        const BigU fix = (BigU::modexp(M.W1, BigU(T), M.N) * M.R1) % M.N;
        if ((rc = set_kconst(ctx, M, K_FIX, fix))) return rc;
        M.fix_T = T + 1;
    }
    const std::string key = "fix:" + std::to_string(M.sh.S);
    Program* P = cached_program(ctx, key);
    if (!P) {
        Program np;
        np.emit(OP_LDI, 0);
        np.emit(OP_MULC, K_FIX); np.n_mul++;
        np.emit(OP_STO, 0);
        if ((rc = program_upload(ctx, np))) return rc;
        P = &(ctx->prog_cache[key] = np);
    }
    IoDesc ins[1] = {{last, (uint32_t)M.sh.S, (uint32_t)M.sh.S}};
    return run_vm(ctx, M, *P, 1, ins, 1, out, M.sh.S, M.sh.S);
}


// signed big integer for the Lagrange coefficients of share combining
struct BigS { BigU mag; bool neg = false; };

// Euclidean division by a small signed integer (Go big.Int.Div semantics, pinned by thresholdkey_test.go:168-177)
BigS euclid_div(const BigS& num, long den) {
    const BigU d((uint64_t)(den < 0 ? -den : den));
    BigU q0, r0; BigU::divmod(num.mag, d, q0, r0);
    BigS q;
    if (!num.neg || num.mag.is_zero()) { q.mag = q0; q.neg = den < 0 && !q0.is_zero(); return q; }
    if (r0.is_zero()) { q.mag = q0; q.neg = den > 0 && !q0.is_zero(); return q; }
    q.mag = q0 + BigU(1); q.neg = den > 0;
    return q;
}

// out[i] = in[i]^-1 mod M; *d_first_bad = index of the first non-invertible item or 0xffffffff
int modinv_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* in, uint32_t* out, uint32_t* d_first_bad) {
    if (M.sh.S > BIG_MAXS) return fail(ctx, PGPU_ERR_UNSUPPORTED, "modulus too wide for modinv");
    CU(ctx, cudaMemsetAsync(d_first_bad, 0xff, 4, ctx->stream));
    InvParams P{(uint32_t)count, M.sh.S, M.d_mod, in, out, d_first_bad};
    CU(ctx, modinv_launch(P, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

// Montgomery's batch inversion: chunks of K items share one extended-Euclid inversion (5 multiplications per item
// instead of one xgcd).  Any non-unit makes its chunk's product a non-unit: the caller then falls back to modinv_dev,
// which names the first offending item.  *all_ok (host) = every chunk product was invertible.
static int batch_inverse_chunks(pgpu_ctx* ctx, const ModCtx& M, size_t chunks, uint32_t K, const uint32_t* in, uint32_t* out,
                                uint32_t* prefix, uint32_t* totals, uint32_t* tinv, uint32_t* d_bad) {
    const uint32_t S = M.sh.S;
    const std::string kf = "binv-f:" + std::to_string(S) + ":" + std::to_string(K), kb = "binv-b:" + std::to_string(S) + ":" + std::to_string(K);
    Program* F = cached_program(ctx, kf);
    int rc;
    if (!F) {
        Program np;
        for (uint32_t k = 0; k < K; ++k) {
            np.emit(OP_LDIO, 0 | (k << 2));
            np.emit(OP_MULC, K_R2); np.n_mul++;                       // x_k * R
            if (k) { np.emit(OP_MULT, 0); np.n_mul++; }              // * prefix_{k-1}
            np.emit(OP_STT, 0); np.use_slot(0);
            np.emit(OP_STOO, 1 | (k << 2));                           // prefix_k (Montgomery form) -> out[1]
        }
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STO, 0);                                           // chunk product, plain
        if ((rc = program_upload(ctx, np))) return rc;
        F = &(ctx->prog_cache[kf] = np);
    }
    Program* B = cached_program(ctx, kb);
    if (!B) {
        Program np;
        np.emit(OP_LDI, 2);                                           // I = (x_0 ... x_{K-1})^-1, plain
        np.emit(OP_STT, 0); np.use_slot(0);
        for (uint32_t k = K - 1; k >= 1; --k) {
            np.emit(OP_MULIO, 1 | ((k - 1) << 2)); np.n_mul++;        // I * prefix_{k-1} * R * R^-1 = x_k^-1
            np.emit(OP_STOO, 0 | (k << 2));
            np.emit(OP_LDIO, 0 | (k << 2));
            np.emit(OP_MULC, K_R2); np.n_mul++;                       // x_k * R
            np.emit(OP_MULT, 0); np.n_mul++;                          // * I * R^-1 = (x_0 ... x_{k-1})^-1
            np.emit(OP_STT, 0);
        }
        np.emit(OP_STOO, 0);
        if ((rc = program_upload(ctx, np))) return rc;
        B = &(ctx->prog_cache[kb] = np);
    }
    IoDesc fin[1] = {{in, K * S, S}};
    if ((rc = run_vm(ctx, M, *F, chunks, fin, 1, totals, S, S, ExpDesc(), prefix, K * S))) return rc;
    InvParams P{(uint32_t)chunks, (int)S, M.d_mod, totals, tinv, d_bad};
    CU(ctx, modinv_launch(P, ctx->stream));
    ctx->launches++;
    IoDesc bin[3] = {{in, K * S, S}, {prefix, K * S, S}, {tinv, S, S}};
    return run_vm(ctx, M, *B, chunks, bin, 3, out, K * S, S);
}

// out[i] = in[i]^-1 mod M for large batches; *d_first_bad as modinv_dev.  in and out may alias.
int modinv_batch_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* in, uint32_t* out, uint32_t* d_first_bad) {
    const uint32_t K = 32, S = M.sh.S;
    if (count < 4 * K || S > BIG_MAXS) return modinv_dev(ctx, M, count, in, out, d_first_bad);
    const size_t chunks = count / K, tail = count % K;
    DEVBUF(prefix, ctx, count * S); DEVBUF(totals, ctx, (chunks + 1) * S); DEVBUF(tinv, ctx, (chunks + 1) * S); DEVBUF(src, ctx, in == out ? count * S : 1);
    const uint32_t* x = in;
    if (in == out) { CU(ctx, cudaMemcpyAsync(src.p, in, count * S * 4, cudaMemcpyDeviceToDevice, ctx->stream)); x = src.p; }
    CU(ctx, cudaMemsetAsync(d_first_bad, 0xff, 4, ctx->stream));
    int rc;
    if ((rc = batch_inverse_chunks(ctx, M, chunks, K, x, out, prefix.p, totals.p, tinv.p, d_first_bad))) return rc;
    if (tail && (rc = batch_inverse_chunks(ctx, M, 1, (uint32_t)tail, x + chunks * K * S, out + chunks * K * S, prefix.p + chunks * K * S,
                                           totals.p + chunks * S, tinv.p + chunks * S, d_first_bad))) return rc;
    uint32_t bad = 0;
    CU(ctx, cudaMemcpyAsync(&bad, d_first_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad != 0xffffffffu) return modinv_dev(ctx, M, count, x, out, d_first_bad);     // a non-unit somewhere: exact per-item pass
    return PGPU_OK;
}

int bigmul_dev(pgpu_ctx* ctx, size_t count, const uint32_t* a, uint32_t na, const uint32_t* b, uint32_t nb, uint32_t* out) {
    MulParams P{(uint32_t)count, a, na, (int)na, b, nb, (int)nb, out, na + nb, na + nb};
    CU(ctx, bigmul_launch(P, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

int sha_dev(pgpu_ctx* ctx, size_t count, int n_seg, const uint32_t* const* seg, const uint32_t* stride, const int* limbs, uint32_t* out,
            const uint32_t* div) {
    ShaParams P{}; P.n_items = (uint32_t)count; P.n_seg = n_seg; P.out = out;
    for (int i = 0; i < n_seg; ++i) { P.seg[i] = seg[i]; P.stride[i] = stride[i]; P.limbs[i] = limbs[i]; P.div[i] = div ? div[i] : 1; }
    CU(ctx, sha256_concat_launch(P, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

// Limbs of a Z record: Z = r + E*delta*share (thresholdkey.go:313-317) with r < n^2, E < 2^256, share < n*m < n^2 and
// delta = l!, i.e. fewer than 2*bitlen(n) + bitlen(l!) + 257 bits.  Never less than the S + 16 limbs of small l (so the
// width of existing records is unchanged up to 57 servers); grows with l! beyond (the reference's tests use 100 servers).
uint32_t z_limbs(const pgpu_ctx* ctx) {
    const uint32_t S = (uint32_t)ctx->m_n2.sh.S;
    const size_t bits = 2 * ctx->n.bitlen() + ctx->tk_delta.bitlen() + 257;
    const uint32_t need = (uint32_t)((bits + 31) / 32);
    return std::max(S + 16, (need + 3) / 4 * 4);
}

// ZKP transcript hash shared by prover and verifier: c^4 and c_i^2 enter unreduced (thresholdkey.go:241,248,319-326)
int zkp_hash_dev(pgpu_ctx* ctx, size_t count, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* dec, uint32_t* e_out) {
    const uint32_t S = ctx->m_n2.sh.S;
    DEVBUF(c2, ctx, count * 2 * S); DEVBUF(c4, ctx, count * 4 * S); DEVBUF(ci2, ctx, count * 2 * S);
    int rc;
    if ((rc = bigmul_dev(ctx, count, c, S, c, S, c2.p))) return rc;
    if ((rc = bigmul_dev(ctx, count, c2.p, 2 * S, c2.p, 2 * S, c4.p))) return rc;
    if ((rc = bigmul_dev(ctx, count, dec, S, dec, S, ci2.p))) return rc;
    const uint32_t* seg[4] = {a, b, c4.p, ci2.p};
    const uint32_t stride[4] = {S, S, 4 * S, 2 * S};
    const int limbs[4] = {(int)S, (int)S, (int)(4 * S), (int)(2 * S)};
    return sha_dev(ctx, count, 4, seg, stride, limbs, e_out);
}

// PartialDecrypt and the proof's a = (c^4)^r in ONE launch (thresholdkey.go:199 and :241-242): both are powers of c --
// c^(2*delta*share) with the key's exponent and c^(4r) with the item's r -- so right to left they share the squaring chain
// c, c^2, c^4, ...: per window of w bits the chain value goes into the bucket of the key exponent's digit (known to the host:
// a static table index) and, two squarings later, into the bucket of r's digit (OP_BKT); the two bucket sets are folded
// into prod_d T[d]^d.  ~6160 squarings + 2*(nwin + 2^(w+1)) multiplications instead of two exponentiations of ~7000 each.
// in0 = c (plain); out[0] = c^e1, out[1] = (c^4)^(item exponent of rbits bits)
void build_pdec_a_program(Program& np, const BigU& e1, uint32_t rbits) {
    const size_t bits = std::max<size_t>(e1.bitlen(), rbits);
    int w = 3;
    { double best = 1e300; for (int cnd = 3; cnd <= 6; ++cnd) { const double cost = 2.0 * ((double)bits / cnd + 2.0 * (1u << cnd)); if (cost < best) { best = cost; w = cnd; } } }
    const uint32_t nb = 1u << w, per = nb + 1, nwin = (uint32_t)((bits + w - 1) / w);
    const uint32_t CH = 2 * per, RUN = CH + 1, ACC = CH + 2;
    np.emit(OP_LDC, K_R1);
    for (uint32_t s = 0; s < 2; ++s) for (uint32_t d = 0; d <= nb; ++d) np.emit(OP_STT, s * per + d);       // every bucket = 1
    np.use_slot(ACC);
    np.emit(OP_LDI, 0);
    np.emit(OP_MULC, K_R2); np.n_mul++;
    for (uint32_t win = 0; win < nwin; ++win) {
        uint32_t d1 = 0;
        for (int b = w - 1; b >= 0; --b) d1 = (d1 << 1) | (e1.bit((size_t)win * w + b) ? 1u : 0u);
        if (d1) {                                                     // chain = c^(2^(w*win)): the key exponent's digit
            np.emit(OP_STT, CH);
            np.emit(OP_MULT, d1); np.n_mul++;
            np.emit(OP_STT, d1);
            np.emit(OP_LDT, CH);
        }
        np.emit(OP_SQR, 2); np.n_sqr += 2;                            // chain = (c^4)^(2^(w*win)): r's digit
        np.emit(OP_STT, CH);
        np.emit(OP_BKT, (win * w) | ((uint32_t)w << 20) | (1u << 24)); np.n_mul++;
        if (win + 1 < nwin) { np.emit(OP_LDT, CH); np.emit(OP_SQR, (uint32_t)w - 2); np.n_sqr += w - 2; }
    }
    for (uint32_t s = 0; s < 2; ++s) {
        np.emit(OP_LDT, s * per + nb - 1);
        np.emit(OP_STT, RUN); np.emit(OP_STT, ACC);
        for (uint32_t d = nb - 2; d >= 1; --d) {
            np.emit(OP_LDT, RUN); np.emit(OP_MULT, s * per + d); np.n_mul++; np.emit(OP_STT, RUN);
            np.emit(OP_MULT, ACC); np.n_mul++; np.emit(OP_STT, ACC);
        }
        np.emit(OP_MULC, K_ONE); np.n_mul++;
        np.emit(OP_STOO, s);                                          // out[0] = c_i, out[1] = a
    }
}

static int pdec_and_a_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, const uint32_t* r, uint32_t* dec, uint32_t* a) {
    ModCtx& M = ctx->m_n2;
    const uint32_t S = M.sh.S;
    const BigU e1 = ctx->tk_share * (BigU(2) * ctx->tk_delta);
    const uint32_t rbits = 32 * S;
    const std::string key = "pdz:" + std::to_string(S) + ":" + e1.hex();
    Program* P = cached_program(ctx, key);
    if (!P) {
        Program np;
        build_pdec_a_program(np, e1, rbits);
        int rc = program_upload(ctx, np);
        if (rc) return rc;
        P = &(ctx->prog_cache[key] = np);
    }
    IoDesc ins[1] = {{c, S, S}};
    ExpDesc ex{r, S, rbits, nullptr, 0};
    return run_vm(ctx, M, *P, count, ins, 1, dec, S, S, ex, a, S);
}

// PartialDecryptionWithZKP (thresholdkey.go:225-255), r supplied
int zkp_prove_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, const uint32_t* r, uint32_t* dec, uint32_t* e, uint32_t* z, bool dec_given) {
    if (!ctx->has_share) return fail(ctx, PGPU_ERR_STATE, "PartialDecryptionWithZKP: no threshold share loaded");
    ModCtx& M = ctx->m_n2;
    const uint32_t S = M.sh.S;
    const BigU k = ctx->tk_delta * ctx->tk_share;
    if (k.v.size() + 8 + 1 > z_limbs(ctx)) return fail(ctx, PGPU_ERR_ARG, "PartialDecryptionWithZKP: the share is not below n^2");
    int rc;
    static const bool no_fused = getenv("PGPU_NO_FUSED_PROVE") != nullptr;
    DEVBUF(c4r, ctx, count * S); DEVBUF(a, ctx, count * S); DEVBUF(b, ctx, count * S); DEVBUF(v, ctx, S);
    DEVBUF(kd, ctx, k.v.size() + 1);
    if ((rc = upload(ctx, v.p, ctx->tk_v.limbs(S)))) return rc;
    if ((rc = upload(ctx, kd.p, k.limbs(k.v.size() + 1)))) return rc;
    if (!dec_given && !no_fused) {
        if ((rc = pdec_and_a_dev(ctx, count, c, r, dec, a.p))) return rc;                // c_i :199 and a = (c^4)^r :242, shared squarings
    } else {
        if (!dec_given && (rc = pdec_dev(ctx, count, c, dec))) return rc;
        if ((rc = modexp_shared_dev(ctx, M, count, c, BigU(4), c4r.p))) return rc;       // c^4 mod n^2: (c^4)^r = (c^4 mod n^2)^r
        if ((rc = modexp_items_dev(ctx, M, count, c4r.p, r, S, a.p))) return rc;         // a = (c^4)^r        :242
    }
    if ((rc = ensure_fix_v(ctx))) return rc;
    if ((rc = modexp_fixed_dev(ctx, M, ctx->fix_v, count, ExpDesc{r, S, 32 * S, nullptr}, b.p))) return rc;   // b = V^r (fixed base) :245
    if ((rc = zkp_hash_dev(ctx, count, a.p, b.p, c, dec, e))) return rc;                 // E                  :250
    MulAddParams Q{(uint32_t)count, r, S, S, e, 8, 8, kd.p, (int)k.v.size(), z, z_limbs(ctx), z_limbs(ctx)};
    CU(ctx, muladd_launch(Q, ctx->stream));                                              // Z = r + E*delta*share :252
    ctx->launches++;
    return PGPU_OK;
}

// VerifyProof (thresholdkey.go:278-311)
// proofs of k servers, n_per_id each, grouped by server: item i belongs to ids[i / n_per_id].  c, dec, e, z, ok hold
// k * n_per_id records in that order.
int zkp_verify_multi_dev(pgpu_ctx* ctx, size_t n_per_id, int k, const int* ids, const uint32_t* c, const uint32_t* dec, const uint32_t* e,
                         const uint32_t* z, uint8_t* ok) {
    if (!ctx->has_threshold) return fail(ctx, PGPU_ERR_STATE, "VerifyProof: no threshold key loaded");
    for (int j = 0; j < k; ++j)
        if (ids[j] < 1 || (size_t)ids[j] > ctx->tk_vi.size())
            return fail(ctx, PGPU_ERR_ARG, "VerifyProof: no verification key for this server id");   // VerificationKeys[ID-1] :305
    const size_t count = n_per_id * (size_t)k;
    if (count == 0) return PGPU_OK;
    ModCtx& M = ctx->m_n2;
    const uint32_t S = M.sh.S, ZL = z_limbs(ctx);
    int rc;
    DEVBUF(t0, ctx, count * S); DEVBUF(t1, ctx, count * S); DEVBUF(t2, ctx, count * S);
    DEVBUF(a, ctx, count * S); DEVBUF(b, ctx, count * S); DEVBUF(kvi, ctx, (size_t)k * S); DEVBUF(bad, ctx, 1); DEVBUF(e2, ctx, count * 8);
    {
        std::vector<uint32_t> vk;
        for (int j = 0; j < k; ++j) { auto l = ctx->tk_vi[ids[j] - 1].limbs(S); vk.insert(vk.end(), l.begin(), l.end()); }
        if ((rc = upload(ctx, kvi.p, vk))) return rc;
    }
    // a = (c^4)^Z * ((c_i^2)^E)^-1 mod n^2        verifyPart1 :293-302
    if ((rc = modexp_shared_dev(ctx, M, count, c, BigU(4), t0.p))) return rc;
    if ((rc = modexp_items_dev(ctx, M, count, t0.p, z, ZL, t1.p))) return rc;
    if ((rc = modmul_dev(ctx, M, count, dec, dec, t0.p))) return rc;
    if ((rc = modexp_items_dev(ctx, M, count, t0.p, e, 8, t2.p))) return rc;
    if ((rc = modinv_batch_dev(ctx, M, count, t2.p, t0.p, bad.p))) return rc;
    if ((rc = modmul_dev(ctx, M, count, t1.p, t0.p, a.p))) return rc;
    // b = V^Z * (v_i^E)^-1 mod n^2                verifyPart2 :304-311
    if ((rc = ensure_fix_v(ctx))) return rc;
    if ((rc = modexp_fixed_dev(ctx, M, ctx->fix_v, count, ExpDesc{z, ZL, 32 * ZL, nullptr}, t1.p))) return rc;   // V^Z (fixed base)
    if ((rc = modexp_items_io(ctx, M, count, IoDesc{kvi.p, S, S, (uint32_t)n_per_id}, ExpDesc{e, 8, 256, nullptr}, t2.p))) return rc;
    if ((rc = modinv_batch_dev(ctx, M, count, t2.p, t0.p, bad.p))) return rc;
    if ((rc = modmul_dev(ctx, M, count, t1.p, t0.p, b.p))) return rc;
    if ((rc = zkp_hash_dev(ctx, count, a.p, b.p, c, dec, e2.p))) return rc;
    CU(ctx, equal_launch(e, e2.p, 8, (uint32_t)count, ok, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

// VerifyProof (thresholdkey.go:278-311) of the proofs of k servers for the SAME n ciphertexts -- what the combiner of a
// threshold round does.  Records are item-major: proof p = i*k + j is server ids[j]'s proof for ciphertext i (c holds n
// records, dec / e / z / ok hold n*k).  The k powers (c^4)^Z of one ciphertext share their squarings (modexp_multi_dev);
// everything else is as in zkp_verify_multi_dev.
int zkp_verify_shared_dev(pgpu_ctx* ctx, size_t n, int k, const int* ids, const uint32_t* c, const uint32_t* dec, const uint32_t* e,
                          const uint32_t* z, uint8_t* ok) {
    if (!ctx->has_threshold) return fail(ctx, PGPU_ERR_STATE, "VerifyProof: no threshold key loaded");
    if (k < 1 || k > 8) return fail(ctx, PGPU_ERR_ARG, "VerifyProof (shared ciphertexts): 1 to 8 servers per call");
    for (int j = 0; j < k; ++j)
        if (ids[j] < 1 || (size_t)ids[j] > ctx->tk_vi.size())
            return fail(ctx, PGPU_ERR_ARG, "VerifyProof: no verification key for this server id");
    const size_t count = n * (size_t)k;
    if (count == 0) return PGPU_OK;
    ModCtx& M = ctx->m_n2;
    const uint32_t S = M.sh.S, ZL = z_limbs(ctx);
    int rc;
    DEVBUF(t0, ctx, count * S); DEVBUF(t1, ctx, count * S); DEVBUF(t2, ctx, count * S);
    DEVBUF(a, ctx, count * S); DEVBUF(b, ctx, count * S); DEVBUF(kvi, ctx, (size_t)k * S); DEVBUF(vrep, ctx, count * S);
    DEVBUF(bad, ctx, 1); DEVBUF(e2, ctx, count * 8);
    {
        std::vector<uint32_t> vk;
        for (int j = 0; j < k; ++j) { auto l = ctx->tk_vi[ids[j] - 1].limbs(S); vk.insert(vk.end(), l.begin(), l.end()); }
        if ((rc = upload(ctx, kvi.p, vk))) return rc;
        CU(ctx, repeat_launch(kvi.p, (uint32_t)k * S, (uint32_t)n, vrep.p, (uint32_t)n, ctx->stream));      // v_i of proof i*k + j = kvi[j]
    }
    // a = (c^4)^Z * ((c_i^2)^E)^-1 mod n^2        verifyPart1 :293-302
    if ((rc = modexp_multi_dev(ctx, M, n, (uint32_t)k, c, 4, z, ZL, t1.p))) return rc;
    if ((rc = modmul_dev(ctx, M, count, dec, dec, t0.p))) return rc;
    if ((rc = modexp_items_dev(ctx, M, count, t0.p, e, 8, t2.p))) return rc;
    if ((rc = modinv_batch_dev(ctx, M, count, t2.p, t0.p, bad.p))) return rc;
    if ((rc = modmul_dev(ctx, M, count, t1.p, t0.p, a.p))) return rc;
    // b = V^Z * (v_i^E)^-1 mod n^2                verifyPart2 :304-311
    if ((rc = ensure_fix_v(ctx))) return rc;
    if ((rc = modexp_fixed_dev(ctx, M, ctx->fix_v, count, ExpDesc{z, ZL, 32 * ZL, nullptr}, t1.p))) return rc;   // V^Z (fixed base)
    if ((rc = modexp_items_dev(ctx, M, count, vrep.p, e, 8, t2.p))) return rc;
    if ((rc = modinv_batch_dev(ctx, M, count, t2.p, t0.p, bad.p))) return rc;
    if ((rc = modmul_dev(ctx, M, count, t1.p, t0.p, b.p))) return rc;
    // E' = SHA-256(a || b || c^4 || c_i^2) with the unreduced c^4 of ciphertext p / k
    {
        DEVBUF(c2, ctx, n * 2 * S); DEVBUF(c4, ctx, n * 4 * S); DEVBUF(ci2, ctx, count * 2 * S);
        if ((rc = bigmul_dev(ctx, n, c, S, c, S, c2.p))) return rc;
        if ((rc = bigmul_dev(ctx, n, c2.p, 2 * S, c2.p, 2 * S, c4.p))) return rc;
        if ((rc = bigmul_dev(ctx, count, dec, S, dec, S, ci2.p))) return rc;
        const uint32_t* seg[4] = {a.p, b.p, c4.p, ci2.p};
        const uint32_t stride[4] = {S, S, 4 * S, 2 * S};
        const int limbs[4] = {(int)S, (int)S, (int)(4 * S), (int)(2 * S)};
        const uint32_t div[4] = {1, 1, (uint32_t)k, 1};
        if ((rc = sha_dev(ctx, count, 4, seg, stride, limbs, e2.p, div))) return rc;
    }
    CU(ctx, equal_launch(e, e2.p, 8, (uint32_t)count, ok, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

// VerifyProof (thresholdkey.go:278-311) for proofs of one server
int zkp_verify_dev(pgpu_ctx* ctx, size_t count, int id, const uint32_t* c, const uint32_t* dec, const uint32_t* e, const uint32_t* z, uint8_t* ok) {
    return zkp_verify_multi_dev(ctx, count, 1, &id, c, dec, e, z, ok);
}

// CombinePartialDecryptions (thresholdkey.go:149-161); decs = k batches of `count` n^2-width records, one per share
int combine_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const uint32_t* decs, uint32_t* m_out, size_t share_stride) {
    if (share_stride == 0) share_stride = count;
    if (!ctx->has_threshold) return fail(ctx, PGPU_ERR_STATE, "CombinePartialDecryptions: no threshold key loaded");
    if (k < ctx->tk_w) return fail(ctx, PGPU_ERR_THRESHOLD, "Threshold not meet");                               // :78-80
    for (int i = 0; i < k; ++i) for (int j = i + 1; j < k; ++j)
        if (ids[i] == ids[j]) return fail(ctx, PGPU_ERR_THRESHOLD, "two shares has been created by the same server");  // :81-87
    ModCtx& M = ctx->m_n2;
    const uint32_t S = M.sh.S;
    const size_t h = ctx->wn;
    if (h > (size_t)CRT_MAXH) return fail(ctx, PGPU_ERR_UNSUPPORTED, "n too wide for the combine tail");
    const BigU Rh = BigU::pow2(32 * h);
    BigU ninv, K;
    if (!BigU::modinv(ctx->n, Rh, ninv)) return fail(ctx, PGPU_ERR_ARG, "n must be odd");
    if (!BigU::modinv((BigU(4) * ctx->tk_delta * ctx->tk_delta) % ctx->n, ctx->n, K)) return fail(ctx, PGPU_ERR_NOT_INVERTIBLE, "4*delta^2 not invertible mod n");
    std::vector<uint32_t> kc;
    for (const BigU& x : {ctx->n, ninv, (K * Rh) % ctx->n}) { auto l = x.limbs(h); kc.insert(kc.end(), l.begin(), l.end()); }
    int rc;
    DEVBUF(pos, ctx, count * S); DEVBUF(neg, ctx, count * S); DEVBUF(t, ctx, count * S); DEVBUF(bad, ctx, 1); DEVBUF(kd, ctx, kc.size());
    if ((rc = upload(ctx, kd.p, kc))) return rc;
    bool have_pos = false, have_neg = false;
    for (int i = 0; i < k; ++i) {
        BigS lam; lam.mag = ctx->tk_delta;                                   // computeLambda :99-107
        for (int j = 0; j < k; ++j) {
            if (ids[j] == ids[i]) continue;
            BigS num; num.mag = lam.mag * BigU((uint64_t)(ids[j] < 0 ? -(long)ids[j] : (long)ids[j]));
            num.neg = (lam.neg != (ids[j] > 0)) && !num.mag.is_zero();       // lambda * (-j)   :92
            lam = euclid_div(num, (long)ids[i] - (long)ids[j]);              // Div(num, i - j) :93-94
        }
        const BigU e2 = lam.mag * BigU(2);                                   // updateCprime: exponent 2*lambda :119-124
        if ((rc = modexp_shared_dev(ctx, M, count, decs + (size_t)i * share_stride * S, e2, t.p))) return rc;
        uint32_t* acc = lam.neg ? neg.p : pos.p;
        bool& have = lam.neg ? have_neg : have_pos;
        if (!have) { CU(ctx, cudaMemcpyAsync(acc, t.p, count * S * 4, cudaMemcpyDeviceToDevice, ctx->stream)); have = true; }
        else if ((rc = modmul_dev(ctx, M, count, acc, t.p, acc))) return rc;
    }
    const uint32_t* cprime = pos.p;
    if (have_neg) {                                                          // negative exponent: ModInverse (exp, :132-138)
        if ((rc = modinv_batch_dev(ctx, M, count, neg.p, t.p, bad.p))) return rc;
        uint32_t first_bad = 0xffffffffu;
        CU(ctx, cudaMemcpyAsync(&first_bad, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (first_bad != 0xffffffffu)           // ModInverse of a non-unit is undefined in the reference (thresholdkey.go:132-138)
            return fail(ctx, PGPU_ERR_NOT_INVERTIBLE, "CombinePartialDecryptions: a partial decryption of item " + std::to_string(first_bad) + " is not a unit mod n^2");
        if (have_pos) { if ((rc = modmul_dev(ctx, M, count, pos.p, t.p, pos.p))) return rc; }
        else cprime = t.p;
    }
    // computeDecryption: L(c') * (4 delta^2)^-1 mod n  :143-146, :63-66
    CombineParams C{(uint32_t)count, (int)h, kd.p, mont_np0(ctx->n.v[0]), cprime, S, m_out};
    CU(ctx, combine_final_launch(C, ctx->stream));
    ctx->launches++;
    return PGPU_OK;
}

// CombinePartialDecryptionsZKP (thresholdkey.go:164-172) for a batch.  The reference filters PER CIPHERTEXT: share j's
// partial decryption of ciphertext i takes part iff its proof verified (ok[j*count + i] != 0).  Ciphertexts are grouped by
// their set of surviving shares -- one CombinePartialDecryptions per distinct set, a single one when every proof holds --
// and a ciphertext left with fewer than `threshold` shares gets a zero plaintext and item_ok = 0 where the reference
// returns "Threshold not meet" for it.  *n_failed counts those.
int combine_verified_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const uint32_t* decs, size_t share_stride, const uint8_t* ok,
                         uint32_t* m_out, uint8_t* item_ok, size_t* n_failed, bool ok_item_major) {
    if (share_stride == 0) share_stride = count;
    if (n_failed) *n_failed = 0;
    if (!ctx->has_threshold) return fail(ctx, PGPU_ERR_STATE, "CombinePartialDecryptionsZKP: no threshold key loaded");
    if (k < 0 || k > 64) return fail(ctx, PGPU_ERR_ARG, "CombinePartialDecryptionsZKP: at most 64 shares per call");
    if (count == 0) return PGPU_OK;
    const uint32_t S = ctx->m_n2.sh.S;
    const size_t h = ctx->wn;
    std::vector<uint8_t> hok((size_t)k * count);
    CU(ctx, cudaMemcpyAsync(hok.data(), ok, hok.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    std::map<uint64_t, std::vector<uint32_t>> groups;
    for (size_t i = 0; i < count; ++i) {
        uint64_t mask = 0;
        for (int j = 0; j < k; ++j) if (hok[ok_item_major ? i * (size_t)k + j : (size_t)j * count + i]) mask |= 1ull << j;
        groups[mask].push_back((uint32_t)i);
    }
    const uint64_t full = k == 64 ? ~0ull : ((1ull << k) - 1);
    std::vector<uint8_t> hitem(count, 1);
    int rc;
    size_t failed = 0;
    if (groups.size() == 1 && groups.begin()->first == full) {
        if ((rc = combine_dev(ctx, count, k, ids, decs, m_out, share_stride))) return rc;
    } else {
        CU(ctx, cudaMemsetAsync(m_out, 0, count * h * 4, ctx->stream));
        for (auto& g : groups) {
            std::vector<int> gids;
            std::vector<int> gj;
            for (int j = 0; j < k; ++j) if (g.first >> j & 1) { gids.push_back(ids[j]); gj.push_back(j); }
            const std::vector<uint32_t>& idx = g.second;
            if ((int)gids.size() < ctx->tk_w) {                                                    // thresholdkey.go:78-80
                for (uint32_t i : idx) hitem[i] = 0;
                failed += idx.size();
                continue;
            }
            const size_t ng = idx.size();
            DEVBUF(didx, ctx, ng); DEVBUF(packed, ctx, gids.size() * ng * S); DEVBUF(mg, ctx, ng * h);
            if ((rc = upload(ctx, didx.p, idx))) return rc;
            for (size_t jj = 0; jj < gj.size(); ++jj)
                CU(ctx, gather_launch(decs + (size_t)gj[jj] * share_stride * S, S, didx.p, 1, packed.p + jj * ng * S, (uint32_t)ng, ctx->stream));
            if ((rc = combine_dev(ctx, ng, (int)gids.size(), gids.data(), packed.p, mg.p, ng))) return rc;
            CU(ctx, scatter_launch(mg.p, (uint32_t)h, didx.p, m_out, (uint32_t)ng, ctx->stream));
            ctx->launches += gj.size() + 1;
        }
    }
    if (item_ok) {
        CU(ctx, cudaMemcpyAsync(item_ok, hitem.data(), count, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));          // hitem leaves scope
    }
    if (n_failed) *n_failed = failed;
    return PGPU_OK;
}

// Encrypted dot product prod c[i]^k[i] (ConstMult + Add, operations.go:11-29,58-64) by Pippenger's bucket method: per
// window of w bits every ciphertext is multiplied into the bucket of its digit (ONE multiplication per item and window,
// no squarings); the buckets are private to a resident group, so the pass is race free and uniform.  Per pass the group
// folds its buckets into prod_d T[d]^d with the running-product trick, the groups' values are multiplied together, and
// the windows are combined by Horner's rule.  ~ (bits/w + 1) multiplications per item instead of bits squarings.
int dot_pippenger_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, const uint32_t* k, uint32_t k_limbs, uint32_t* out) {
    ModCtx& M = ctx->m_n2;
    const uint32_t S = M.sh.S, w = 4, nb = 1u << w, bits = 32 * k_limbs, nwin = (bits + w - 1) / w;
    const int blocks = vm_full_blocks(ctx, M);
    const int gpb = VM_BLOCK_THREADS / M.sh.tpi;
    const size_t n_groups = (size_t)blocks * gpb;
    int rc;
    auto program = [&](const std::string& key, auto&& build) -> Program* {
        Program* P = cached_program(ctx, key);
        if (P) return P;
        Program np;
        build(np);
        if (program_upload(ctx, np)) return nullptr;
        return &(ctx->prog_cache[key] = np);
    };
    Program* p_tom = program("tom:" + std::to_string(S), [&](Program& p) {
        p.emit(OP_LDI, 0); p.emit(OP_MULC, K_R2); p.n_mul++; p.emit(OP_STO, 0);
    });
    Program* p_init = program("bkt-init:" + std::to_string(S), [&](Program& p) {
        p.emit(OP_LDC, K_R1);
        for (uint32_t d = 0; d <= nb; ++d) p.emit(OP_STT, d);
        p.use_slot(nb + 2);      // same table layout as the two programs below: the buckets must survive between launches
    });
    Program* p_agg = program("bkt-agg:" + std::to_string(S), [&](Program& p) {
        // run = T[nb-1]; acc = run; for d = nb-2 .. 1: run *= T[d]; acc *= run    (acc = prod_d T[d]^d)
        const uint32_t RUN = nb + 1, ACC = nb + 2;
        p.use_slot(ACC);
        p.emit(OP_LDT, nb - 1); p.emit(OP_STT, RUN); p.emit(OP_STT, ACC);
        for (uint32_t d = nb - 2; d >= 1; --d) {
            p.emit(OP_LDT, RUN); p.emit(OP_MULT, d); p.n_mul++; p.emit(OP_STT, RUN);
            p.emit(OP_MULT, ACC); p.n_mul++; p.emit(OP_STT, ACC);
        }
        p.emit(OP_MULC, K_ONE); p.n_mul++;
        p.emit(OP_STO, 0);
    });
    if (!p_tom || !p_init || !p_agg) return fail(ctx, PGPU_ERR_CUDA, "dot product: program upload failed");
    DEVBUF(cM, ctx, count * S); DEVBUF(accs, ctx, n_groups * S); DEVBUF(W, ctx, (size_t)nwin * S); DEVBUF(res, ctx, 2 * S);
    IoDesc in_c[1] = {{c, S, S}};
    if ((rc = run_vm(ctx, M, *p_tom, count, in_c, 1, cM.p, S, S))) return rc;                     // c -> Montgomery form, once
    for (uint32_t j = 0; j < nwin; ++j) {
        Program* p_bkt = program("bkt:" + std::to_string(S) + ":" + std::to_string(j * w), [&](Program& p) {
            p.emit(OP_LDI, 0);
            p.emit(OP_BKT, (j * w) | (w << 20)); p.n_mul++;
            p.use_slot(nb + 2);
        });
        if (!p_bkt) return fail(ctx, PGPU_ERR_CUDA, "dot product: program upload failed");
        // all three launches use the same resident grid: the buckets live in the groups' table slots
        if ((rc = run_vm(ctx, M, *p_init, n_groups, nullptr, 0, accs.p, S, S, ExpDesc(), nullptr, 0, blocks))) return rc;
        IoDesc in_m[1] = {{cM.p, S, S}};
        if ((rc = run_vm(ctx, M, *p_bkt, count, in_m, 1, accs.p, S, S, ExpDesc{k, k_limbs, bits, nullptr}, nullptr, 0, blocks))) return rc;
        if ((rc = run_vm(ctx, M, *p_agg, n_groups, nullptr, 0, accs.p, S, S, ExpDesc(), nullptr, 0, blocks))) return rc;
        if ((rc = prod_dev(ctx, M, n_groups, accs.p, W.p + (size_t)j * S))) return rc;               // window value, plain
    }
    // Horner over the windows: res = (...(W[nwin-1]^(2^w) * W[nwin-2])^(2^w) ... ) * W[0]
    CU(ctx, cudaMemcpyAsync(res.p, W.p + (size_t)(nwin - 1) * S, S * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    for (uint32_t j = nwin - 1; j-- > 0;) {
        if ((rc = modexp_shared_dev(ctx, M, 1, res.p, BigU((uint64_t)nb), res.p + S))) return rc;
        if ((rc = modmul_dev(ctx, M, 1, res.p + S, W.p + (size_t)j * S, res.p))) return rc;
    }
    CU(ctx, cudaMemcpyAsync(out, res.p, S * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return PGPU_OK;
}

uint32_t* HostIo::in(int slot, const void* host, size_t bytes) {
    void* d = nullptr;
    if (rc) return nullptr;
    if ((rc = stage(ctx, slot, bytes, &d))) return nullptr;
    cudaError_t e = cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { rc = fail(ctx, PGPU_ERR_CUDA, std::string("H2D copy: ") + cudaGetErrorString(e)); return nullptr; }
    return (uint32_t*)d;
}
uint32_t* HostIo::out(int slot, size_t bytes) {
    void* d = nullptr;
    if (rc) return nullptr;
    if ((rc = stage(ctx, slot, bytes, &d))) return nullptr;
    return (uint32_t*)d;
}
int HostIo::finish(void* host, const uint32_t* dev, size_t bytes) {
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, PGPU_ERR_CUDA, std::string("D2H copy: ") + cudaGetErrorString(e));
    return PGPU_OK;
}

int set_device(pgpu_ctx* ctx) { CU(ctx, cudaSetDevice(ctx->device)); return PGPU_OK; }

size_t resident_groups(const pgpu_ctx* ctx, const ModCtx& m) {
    return (size_t)ctx->sms * (size_t)std::max(m.blocks_per_sm, 1) * (VM_BLOCK_THREADS / std::max(m.sh.tpi, 1));
}

// Items per chunk: whole rounds of the persistent grid (`align` resident groups), about 2^15 items (16 MB of 512-byte
// records: large enough for the full PCIe rate, small enough that a 2^16 batch of a light operation already overlaps its
// copies); batches of up to two grids go in one piece.  PGPU_CHUNK_ITEMS overrides (0 = never chunk).
size_t chunk_items(size_t count, size_t align) {
    static const long forced = [] { const char* e = getenv("PGPU_CHUNK_ITEMS"); return e ? atol(e) : -1L; }();
    if (forced == 0) return count;
    if (forced > 0) return std::min((size_t)forced, count);        // tests: exact size, ragged against the grid
    align = std::max<size_t>(align, 1);
    if (count <= 2 * align) return count;
    const size_t target = (size_t)1 << 15;
    const size_t chunk = ((target + align - 1) / align) * align;
    return std::min(chunk, count);
}

static int chunk_streams(pgpu_ctx* ctx) {
    if (ctx->s_in) return PGPU_OK;
    CU(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CU(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_in[b], cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_cmp[b], cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_out[b], cudaEventDisableTiming));
    }
    return PGPU_OK;
}

static int run_chunked_impl(pgpu_ctx* ctx, size_t count, const ChunkedIo& io,
                            const std::function<int(size_t, const uint32_t* const*, uint32_t* const*)>& fn);

int run_chunked(pgpu_ctx* ctx, size_t count, const ChunkedIo& io,
                const std::function<int(size_t, const uint32_t* const*, uint32_t* const*)>& fn) {
    if (count == 0) return PGPU_OK;
    int rc;
    if ((rc = chunk_streams(ctx))) return rc;
    rc = run_chunked_impl(ctx, count, io, fn);
    if (rc) {
        // whatever failed, no copy may still be reading or writing the caller's buffers when the call returns
        cudaStreamSynchronize(ctx->s_in); cudaStreamSynchronize(ctx->s_out); cudaStreamSynchronize(ctx->stream);
    }
    return rc;
}

static int run_chunked_impl(pgpu_ctx* ctx, size_t count, const ChunkedIo& io,
                            const std::function<int(size_t, const uint32_t* const*, uint32_t* const*)>& fn) {
    int rc;
    const size_t chunk = chunk_items(count, io.align);
    const size_t n_chunks = (count + chunk - 1) / chunk;
    const int n_buf = n_chunks > 1 ? 2 : 1;
    // staging: slot b*5 + i for input i, b*5 + 3 + o for output o of buffer set b
    uint32_t* din[2][ChunkedIo::MAX_IN] = {};
    uint32_t* dout[2][ChunkedIo::MAX_OUT] = {};
    for (int b = 0; b < n_buf; ++b) {
        for (int i = 0; i < io.n_in; ++i) { void* d; if ((rc = stage(ctx, b * 5 + i, chunk * io.in_w[i], &d))) return rc; din[b][i] = (uint32_t*)d; }
        for (int o = 0; o < io.n_out; ++o) { void* d; if ((rc = stage(ctx, b * 5 + 3 + o, chunk * io.out_w[o], &d))) return rc; dout[b][o] = (uint32_t*)d; }
    }
    auto items_of = [&](size_t k) { return std::min(chunk, count - k * chunk); };
    auto h2d = [&](size_t k) -> int {
        const int b = (int)(k & 1);
        CU(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_cmp[b], 0));      // chunk k-2 (or whatever ran before this call) has consumed these inputs
        for (int i = 0; i < io.n_in; ++i)
            CU(ctx, cudaMemcpyAsync(din[b][i], (const uint8_t*)io.in[i] + k * chunk * io.in_w[i], items_of(k) * io.in_w[i], cudaMemcpyHostToDevice, ctx->s_in));
        CU(ctx, cudaEventRecord(ctx->ev_in[b], ctx->s_in));
        return PGPU_OK;
    };
    // anything already enqueued on the context's stream (a previous call's kernels reading the staging slots) comes first
    CU(ctx, cudaEventRecord(ctx->ev_cmp[0], ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev_cmp[1], ctx->stream));
    CU(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_cmp[0], 0));
    if ((rc = h2d(0))) return rc;
    if (ctx->timing) cudaEventRecord(ctx->ev0, ctx->stream);
    for (size_t k = 0; k < n_chunks; ++k) {
        const int b = (int)(k & 1);
        CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
        if (k >= 2) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[b], 0));          // chunk k-2's results have left
        if ((rc = fn(items_of(k), din[b], dout[b]))) return rc;
        CU(ctx, cudaEventRecord(ctx->ev_cmp[b], ctx->stream));
        if (k + 1 < n_chunks && (rc = h2d(k + 1))) return rc;
        CU(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_cmp[b], 0));
        for (int o = 0; o < io.n_out; ++o)
            CU(ctx, cudaMemcpyAsync((uint8_t*)io.out[o] + k * chunk * io.out_w[o], dout[b][o], items_of(k) * io.out_w[o], cudaMemcpyDeviceToHost, ctx->s_out));
        CU(ctx, cudaEventRecord(ctx->ev_out[b], ctx->s_out));
    }
    if (ctx->timing) { cudaEventRecord(ctx->ev1, ctx->stream); ctx->ev_valid = true; }
    CU(ctx, cudaStreamSynchronize(ctx->s_out));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return PGPU_OK;
}

}  // namespace pgpu
