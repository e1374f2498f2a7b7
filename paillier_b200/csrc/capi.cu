// C ABI of libpaillier_b200.so (include/pgpu.h): argument checking, host-buffer staging and
// dispatch into the engine (engine.hpp).  No CPU arithmetic happens on a batch item here.
#include "engine.hpp"

using namespace pgpu;

#define g_err (::pgpu::thread_error())

// =========================================================== extern "C" API
// Host interpreter of the micro-programs (vm.h) on BigU integers, for the CPU test suite only (tests/test_host_programs.py):
// the exponentiation schedules the host compiles -- sliding window over a shared exponent, fixed windows over per-item exponents,
// the right-to-left bucket programs of the proofs -- are executed with plain modular arithmetic (Montgomery radix 1) and compared
// with pow().  It checks the COMPILER of the programs without a GPU; no entry point of the product routes through it.
namespace {
uint32_t host_exp_bits(const uint32_t* e, uint32_t nbits, uint32_t pos, uint32_t w) {
    if (pos >= nbits) return 0;
    const uint32_t limb = pos >> 5, sh = pos & 31, nlimbs = (nbits + 31) >> 5;
    uint64_t v = e[limb];
    if (sh + w > 32 && limb + 1 < nlimbs) v |= (uint64_t)e[limb + 1] << 32;
    uint32_t r = (uint32_t)(v >> sh) & ((1u << w) - 1u);
    if (pos + w > nbits) r &= (1u << (nbits - pos)) - 1u;
    return r;
}

int host_vm_run(const Program& P, const BigU& N, const BigU& in0, const uint32_t* exps, uint32_t exp_limbs, uint32_t exp_sub, std::vector<BigU>& outs) {
    const uint32_t nbits = 32 * exp_limbs;
    std::map<uint32_t, BigU> T;
    BigU x;
    auto mulmod = [&](const BigU& a, const BigU& b) { return (a * b) % N; };
    auto kconst = [&](uint32_t slot, BigU& v) { if (slot == K_R1 || slot == K_R2 || slot == K_ONE) { v = BigU(1); return true; } return false; };
    for (uint32_t op : P.ops) {
        const uint32_t code = op >> 27, arg = op & 0x07ffffffu;
        uint32_t nsq = 0; bool mul = false; BigU y; uint32_t bkt = 0xffffffffu;
        switch (code) {
            case OP_END: return PGPU_OK;
            case OP_LDI: if (arg != 0) return PGPU_ERR_UNSUPPORTED; x = in0; break;
            case OP_LDC: if (!kconst(arg, x)) return PGPU_ERR_UNSUPPORTED; break;
            case OP_LDT: x = T[arg]; break;
            case OP_STT: T[arg] = x; break;
            case OP_STO: if (outs.size() < 1) outs.resize(1); outs[0] = x % N; break;
            case OP_STOO: { const uint32_t slot = (arg & 1u) ? 1u : (arg >> 2);        // out[1][item] or out[0][item, off]
                            if (outs.size() <= slot) outs.resize(slot + 1); outs[slot] = x % N; } break;
            case OP_SQR: nsq = arg; break;
            case OP_MULT: y = T[arg]; mul = true; break;
            case OP_MULC: if (!kconst(arg, y)) return PGPU_ERR_UNSUPPORTED; mul = true; break;
            case OP_WIN: { const uint32_t pos = arg & 0xfffffu, w = (arg >> 20) & 0xfu, tbase = arg >> 24;
                           y = T[tbase + host_exp_bits(exps, nbits, pos, w)]; nsq = w; mul = true; } break;
            case OP_SQMT: y = T[arg >> 12]; nsq = arg & 0xfffu; mul = true; break;
            case OP_BKT: { const uint32_t pos = arg & 0xfffffu, w = (arg >> 20) & 0xfu, sub = arg >> 24;
                           bkt = sub * ((1u << w) + 1u) + host_exp_bits(exps + (size_t)sub * exp_sub, nbits, pos, w);      // VmParams::exp_sub
                           y = T[bkt]; mul = true; } break;
            default: return PGPU_ERR_UNSUPPORTED;
        }
        for (uint32_t i = 0; i < nsq; ++i) x = mulmod(x, x);
        if (mul) x = mulmod(x, y);
        if (bkt != 0xffffffffu) T[bkt] = x;
    }
    return PGPU_OK;
}
}  // namespace


namespace {
// the program pgpu_selftest_program compiled last on this thread (pgpu_selftest_last_program hands it to the CPU test that runs
// the interpreter's source on an emulated warp)
std::vector<uint32_t>& selftest_last_ops() { static thread_local std::vector<uint32_t> v; return v; }
uint32_t& selftest_last_tbl() { static thread_local uint32_t t = 0; return t; }
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int pgpu_version(void) { return 1; }

int pgpu_device_count(int* count) {
    if (!count) return fail(nullptr, PGPU_ERR_ARG, "null pointer");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return fail(nullptr, PGPU_ERR_CUDA, cudaGetErrorString(e)); }
    return PGPU_OK;
}

const char* pgpu_last_error(const pgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int pgpu_ctx_create(pgpu_ctx** out, int device, const uint8_t* n_be, size_t n_len) {
    GUARD_BEGIN
    REQUIRE(nullptr, out && n_be && n_len > 0, "pgpu_ctx_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PGPU_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0"));
    REQUIRE(nullptr, device >= 0 && device < ndev, "pgpu_ctx_create: bad device index");
    pgpu_ctx* ctx = new pgpu_ctx();
    ctx->device = device;
    int rc = PGPU_OK;
    auto bail = [&](int code) { std::string m = ctx->err; pgpu_ctx_destroy(ctx); g_err = m; return code; };
    if ((rc = set_device(ctx))) return bail(rc);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(fail(ctx, PGPU_ERR_CUDA, "cudaGetDeviceProperties failed"));
    ctx->sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(ctx, PGPU_ERR_CUDA, "cudaStreamCreate failed"));
    ctx->stream = ctx->own_stream;
    {   // keep the stream-ordered pool warm: the temporaries of the composite operations come from it
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
    ctx->n = BigU::from_be(n_be, n_len);
    if (!ctx->n.is_odd() || ctx->n.bitlen() < 2) return bail(fail(ctx, PGPU_ERR_ARG, "n must be an odd integer >= 3"));
    ctx->n2 = ctx->n * ctx->n;
    ctx->n3 = ctx->n2 * ctx->n;
    if ((rc = modctx_init(ctx, ctx->m_n2, ctx->n2))) return bail(rc);
    ctx->wn = ctx->m_n2.sh.S / 2;
    if (ctx->n.v.size() > ctx->wn) return bail(fail(ctx, PGPU_ERR_UNSUPPORTED, "n does not fit half an n^2 record"));
    if ((rc = modctx_init(ctx, ctx->m_n, ctx->n))) return bail(rc);
    {   // level 2 is optional: n^3 may exceed the built shapes
        Shape s3;
        if (pick_shape(ctx->n3.v.size(), s3)) { if ((rc = modctx_init(ctx, ctx->m_n3, ctx->n3))) return bail(rc); }
    }
    // n * R^2 mod n^2 for the g = n+1 shortcut
    if ((rc = set_kconst(ctx, ctx->m_n2, K_NR2, (ctx->n * ctx->m_n2.R2) % ctx->n2))) return bail(rc);
    if ((rc = build_encrypt(ctx))) return bail(rc);
    if ((rc = setup_level2(ctx))) return bail(rc);
    *out = ctx;
    return PGPU_OK;
    GUARD_END(nullptr)
}

int pgpu_ctx_destroy(pgpu_ctx* ctx) {
    if (!ctx) return PGPU_OK;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    modctx_free(ctx->m_n); modctx_free(ctx->m_n2); modctx_free(ctx->m_n3); modctx_free(ctx->m_p2); modctx_free(ctx->m_q2);
    program_free(ctx->prog_enc); program_free(ctx->prog_dec_p); program_free(ctx->prog_dec_q); program_free(ctx->prog_pdec);
    program_free(ctx->prog_encq); program_free(ctx->prog_encp); program_free(ctx->prog_encf);
    for (auto& kv : ctx->prog_cache) { dev_scrub_free(kv.second.d_ops, kv.second.d_bytes); scrub(kv.second.ops); }
    protocols_free(ctx);
    // key material and plaintext scratch do not go back to the allocator as they are (engine.hpp: dev_scrub_free)
    dev_scrub_free(ctx->d_crt, ctx->d_crt_bytes);
    dev_scrub_free(ctx->d_table, ctx->table_limbs * 4);
    for (int i = 0; i < 16; ++i) dev_scrub_free(ctx->d_stage[i], ctx->stage_bytes[i]);
    scrub(ctx->p); scrub(ctx->q); scrub(ctx->tk_share);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->s_in) { cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_out); }
    for (int b = 0; b < 2; ++b) {
        if (ctx->ev_in[b]) cudaEventDestroy(ctx->ev_in[b]);
        if (ctx->ev_cmp[b]) cudaEventDestroy(ctx->ev_cmp[b]);
        if (ctx->ev_out[b]) cudaEventDestroy(ctx->ev_out[b]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return PGPU_OK;
}

int pgpu_ctx_widths(const pgpu_ctx* ctx, size_t* w_n, size_t* w_n2, size_t* w_n3) {
    if (!ctx) return fail(nullptr, PGPU_ERR_ARG, "null context");
    if (w_n) *w_n = ctx->wn * 4;
    if (w_n2) *w_n2 = (size_t)ctx->m_n2.sh.S * 4;
    if (w_n3) *w_n3 = ctx->m_n3.ready ? (size_t)ctx->m_n3.sh.S * 4 : 0;
    return PGPU_OK;
}

int pgpu_ctx_mod_width(const pgpu_ctx* ctx, int modsel, size_t* width) {
    if (!ctx || !width) return fail(nullptr, PGPU_ERR_ARG, "null argument");
    const ModCtx* M = modsel == PGPU_MOD_N ? &ctx->m_n : modsel == PGPU_MOD_N2 ? &ctx->m_n2 : modsel == PGPU_MOD_N3 ? &ctx->m_n3 : nullptr;
    *width = (M && M->ready) ? (size_t)M->sh.S * 4 : 0;
    return PGPU_OK;
}

int pgpu_ctx_set_stream(pgpu_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(nullptr, PGPU_ERR_ARG, "null context");
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PGPU_OK;
}

int pgpu_ctx_set_secret_pq(pgpu_ctx* ctx, const uint8_t* p_be, size_t p_len, const uint8_t* q_be, size_t q_len) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && p_be && q_be, "pgpu_ctx_set_secret_pq: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    ctx->p = BigU::from_be(p_be, p_len); ctx->q = BigU::from_be(q_be, q_len);
    REQUIRE(ctx, ctx->p * ctx->q == ctx->n, "p*q != n");
    REQUIRE(ctx, ctx->p != ctx->q, "p == q");
    return setup_crt(ctx);
    GUARD_END(ctx)
}

int pgpu_ctx_set_secret_lambda(pgpu_ctx* ctx, const uint8_t* lambda_be, size_t lambda_len) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && lambda_be, "pgpu_ctx_set_secret_lambda: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    // lambda = (p-1)(q-1) = n - (p+q) + 1  =>  s = p+q = n - lambda + 1, (p-q)^2 = s^2 - 4n
    const BigU lambda = BigU::from_be(lambda_be, lambda_len);
    REQUIRE(ctx, lambda < ctx->n, "lambda >= n");
    const BigU s = ctx->n - lambda + BigU(1);
    const BigU s2 = s * s, n4 = ctx->n * BigU(4);
    REQUIRE(ctx, s2 >= n4, "lambda is not (p-1)(q-1) for this n");
    const BigU d = BigU::isqrt(s2 - n4);
    REQUIRE(ctx, d * d == s2 - n4, "lambda is not (p-1)(q-1) for this n");
    ctx->p = (s - d).shr(1); ctx->q = (s + d).shr(1);
    REQUIRE(ctx, ctx->p * ctx->q == ctx->n, "could not recover p, q from lambda");
    REQUIRE(ctx, ctx->p != ctx->q, "p == q");
    return setup_crt(ctx);
    GUARD_END(ctx)
}

int pgpu_ctx_set_threshold(pgpu_ctx* ctx, int total_servers, int threshold, int id,
                           const uint8_t* share_be, size_t share_len, const uint8_t* v_be, size_t v_len,
                           const void* vkeys_n2w) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx != nullptr, "null context");
    REQUIRE(ctx, total_servers >= 1 && total_servers <= 1000, "bad TotalNumberOfDecryptionServers");
    REQUIRE(ctx, threshold >= 1 && threshold <= total_servers, "Threshold must be between 1 and TotalNumberOfDecryptionServers");
    REQUIRE(ctx, !share_be || (id >= 1 && id <= total_servers), "the share-holder's id must be between 1 and TotalNumberOfDecryptionServers");
    REQUIRE(ctx, id >= 0 && id <= total_servers, "bad server id");
    int rc; if ((rc = set_device(ctx))) return rc;
    ctx->tk_l = total_servers; ctx->tk_w = threshold; ctx->tk_id = id;
    ctx->tk_delta = BigU::factorial((unsigned)total_servers);
    ctx->tk_v = v_be ? BigU::from_be(v_be, v_len) : BigU();
    cudaStreamSynchronize(ctx->stream);
    fixed_table_free(ctx->fix_v);
    ctx->tk_vi.clear();
    if (vkeys_n2w) {
        const uint32_t* p = (const uint32_t*)vkeys_n2w;
        for (int i = 0; i < total_servers; ++i) ctx->tk_vi.push_back(BigU::from_limbs(p + (size_t)i * ctx->m_n2.sh.S, ctx->m_n2.sh.S));
    }
    ctx->has_threshold = true;
    ctx->has_share = false;
    if (share_be) {
        ctx->tk_share = BigU::from_be(share_be, share_len);
        if ((rc = build_pdec(ctx))) return rc;
        ctx->has_share = true;
    }
    return PGPU_OK;
    GUARD_END(ctx)
}

// ---- device-pointer entry points
int pgpu_encrypt_with_r_dev(pgpu_ctx* ctx, size_t count, const void* m, const void* r, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && r && c)), "pgpu_encrypt_with_r: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return encrypt_dev(ctx, count, (const uint32_t*)m, (const uint32_t*)r, (uint32_t*)c);
    GUARD_END(ctx)
}

int pgpu_encrypt_with_r_sk_dev(pgpu_ctx* ctx, size_t count, const void* m, const void* r, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && r && c)), "pgpu_encrypt_with_r_sk: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return encrypt_crt_dev(ctx, count, (const uint32_t*)m, (const uint32_t*)r, (uint32_t*)c);
    GUARD_END(ctx)
}

int pgpu_encrypt_with_rn_dev(pgpu_ctx* ctx, size_t count, const void* m, const void* rn, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && rn && c)), "pgpu_encrypt_with_rn: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return encrypt_rn_dev(ctx, count, (const uint32_t*)m, (const uint32_t*)rn, (uint32_t*)c);
    GUARD_END(ctx)
}

int pgpu_decrypt_dev(pgpu_ctx* ctx, size_t count, const void* c, void* m) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && c)), "pgpu_decrypt: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return decrypt_dev(ctx, count, (const uint32_t*)c, (uint32_t*)m);
    GUARD_END(ctx)
}

int pgpu_partial_decrypt_dev(pgpu_ctx* ctx, size_t count, const void* c, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (out && c)), "pgpu_partial_decrypt: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return pdec_dev(ctx, count, (const uint32_t*)c, (uint32_t*)out);
    GUARD_END(ctx)
}

int pgpu_const_mult_dev(pgpu_ctx* ctx, size_t count, const void* c, const void* k, size_t k_bytes, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && k && out)), "pgpu_const_mult: null argument");
    REQUIRE(ctx, k_bytes > 0 && k_bytes % 4 == 0, "pgpu_const_mult: k_bytes must be a positive multiple of 4");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return modexp_items_dev(ctx, ctx->m_n2, count, (const uint32_t*)c, (const uint32_t*)k, (uint32_t)(k_bytes / 4), (uint32_t*)out);
    GUARD_END(ctx)
}

int pgpu_add_reduce_dev(pgpu_ctx* ctx, size_t count, const void* c, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && out && (count == 0 || c), "pgpu_add_reduce: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return prod_dev(ctx, ctx->m_n2, count, (const uint32_t*)c, (uint32_t*)out);
    GUARD_END(ctx)
}

int pgpu_dot_u64_dev(pgpu_ctx* ctx, size_t count, const void* c, const uint64_t* k, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && out && (count == 0 || (c && k)), "pgpu_dot_u64: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    // large batches: Pippenger's bucket method (17 multiplications per term instead of ~94); small ones: ConstMult + Add
    const size_t resident = (size_t)vm_full_blocks(ctx, ctx->m_n2) * (VM_BLOCK_THREADS / ctx->m_n2.sh.tpi);
    if (count >= 16 * resident && !getenv("PGPU_DOT_SIMPLE"))
        return dot_pippenger_dev(ctx, count, (const uint32_t*)c, (const uint32_t*)k, 2, (uint32_t*)out);
    void* tmp;
    if ((rc = stage(ctx, 15, std::max<size_t>(count, 1) * ctx->m_n2.sh.S * 4, &tmp))) return rc;
    if ((rc = modexp_items_dev(ctx, ctx->m_n2, count, (const uint32_t*)c, (const uint32_t*)k, 2, (uint32_t*)tmp))) return rc;
    return prod_dev(ctx, ctx->m_n2, count, (const uint32_t*)tmp, (uint32_t*)out);
    GUARD_END(ctx)
}

int pgpu_add_pairs_dev(pgpu_ctx* ctx, size_t count, const void* a, const void* b, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (a && b && out)), "pgpu_add_pairs_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return modmul_dev(ctx, ctx->m_n2, count, (const uint32_t*)a, (const uint32_t*)b, (uint32_t*)out);
    GUARD_END(ctx)
}

/* first_bad: device word, 0xffffffff when every b[i] is a unit, else the index of the first one that is not */
int pgpu_sub_pairs_dev(pgpu_ctx* ctx, size_t count, const void* a, const void* b, void* out, uint32_t* first_bad) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && first_bad && (count == 0 || (a && b && out)), "pgpu_sub_pairs_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    DEVBUF(inv, ctx, count * ctx->m_n2.sh.S);        // out may alias a or b
    if ((rc = modinv_batch_dev(ctx, ctx->m_n2, count, (const uint32_t*)b, inv.p, first_bad))) return rc;
    return modmul_dev(ctx, ctx->m_n2, count, (const uint32_t*)a, inv.p, (uint32_t*)out);
    GUARD_END(ctx)
}

int pgpu_randomize_with_r_dev(pgpu_ctx* ctx, size_t count, const void* c, const void* r, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && r && out)), "pgpu_randomize_with_r_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return randomize_dev(ctx, count, (const uint32_t*)c, (const uint32_t*)r, (uint32_t*)out);
    GUARD_END(ctx)
}

// ---- device buffers, pinned host memory (SURVEY.md 8b "Ownership", 8f rank 1)
int pgpu_buf_alloc(pgpu_ctx* ctx, size_t bytes, pgpu_buf** out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && out, "pgpu_buf_alloc: null argument");
    *out = nullptr;
    int rc; if ((rc = set_device(ctx))) return rc;
    pgpu_buf* b = new pgpu_buf();
    b->ctx = ctx; b->bytes = bytes;
    cudaError_t e = cudaMalloc(&b->ptr, std::max<size_t>(bytes, 1));
    if (e != cudaSuccess) { delete b; return fail(ctx, PGPU_ERR_CUDA, std::string("pgpu_buf_alloc: ") + cudaGetErrorString(e)); }
    *out = b;
    return PGPU_OK;
    GUARD_END(ctx)
}

int pgpu_buf_free(pgpu_buf* b) {
    if (!b) return PGPU_OK;
    pgpu_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);            // work enqueued on the buffer has finished
    cudaFree(b->ptr);
    delete b;
    return PGPU_OK;
}

void* pgpu_buf_ptr(const pgpu_buf* b) { return b ? b->ptr : nullptr; }
size_t pgpu_buf_size(const pgpu_buf* b) { return b ? b->bytes : 0; }

int pgpu_buf_upload(pgpu_buf* dst, size_t dst_off, const void* host, size_t bytes) {
    if (!dst) return fail(nullptr, PGPU_ERR_ARG, "pgpu_buf_upload: null buffer");
    pgpu_ctx* ctx = dst->ctx;
    GUARD_BEGIN
    REQUIRE(ctx, bytes == 0 || host, "pgpu_buf_upload: null host pointer");
    REQUIRE(ctx, dst_off <= dst->bytes && bytes <= dst->bytes - dst_off, "pgpu_buf_upload: range outside the buffer");
    int rc; if ((rc = set_device(ctx))) return rc;
    CU(ctx, cudaMemcpyAsync((uint8_t*)dst->ptr + dst_off, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));   // the library keeps no host pointer past return
    return PGPU_OK;
    GUARD_END(ctx)
}

int pgpu_buf_download(const pgpu_buf* src, size_t src_off, void* host, size_t bytes) {
    if (!src) return fail(nullptr, PGPU_ERR_ARG, "pgpu_buf_download: null buffer");
    pgpu_ctx* ctx = src->ctx;
    GUARD_BEGIN
    REQUIRE(ctx, bytes == 0 || host, "pgpu_buf_download: null host pointer");
    REQUIRE(ctx, src_off <= src->bytes && bytes <= src->bytes - src_off, "pgpu_buf_download: range outside the buffer");
    int rc; if ((rc = set_device(ctx))) return rc;
    CU(ctx, cudaMemcpyAsync(host, (const uint8_t*)src->ptr + src_off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return PGPU_OK;
    GUARD_END(ctx)
}

int pgpu_ctx_sync(pgpu_ctx* ctx) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx, "pgpu_ctx_sync: null context");
    int rc; if ((rc = set_device(ctx))) return rc;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return PGPU_OK;
    GUARD_END(ctx)
}

int pgpu_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(nullptr, PGPU_ERR_ARG, "pgpu_host_alloc: null argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(nullptr, PGPU_ERR_CUDA, std::string("pgpu_host_alloc: ") + cudaGetErrorString(e));
    return PGPU_OK;
}

int pgpu_host_free(void* p) {
    if (!p) return PGPU_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) return fail(nullptr, PGPU_ERR_CUDA, std::string("pgpu_host_free: ") + cudaGetErrorString(e));
    return PGPU_OK;
}

// ---- host-buffer entry points
int pgpu_encrypt_with_r(pgpu_ctx* ctx, size_t count, const void* m, const void* r, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && r && c)), "pgpu_encrypt_with_r: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    const size_t wn = ctx->wn * 4, w2 = (size_t)ctx->m_n2.sh.S * 4;
    ChunkedIo io;
    io.add_in(m, wn); io.add_in(r, wn); io.add_out(c, w2);
    io.align = resident_groups(ctx, ctx->m_n2);
    return run_chunked(ctx, count, io, [&](size_t n, const uint32_t* const* in, uint32_t* const* out) { return encrypt_dev(ctx, n, in[0], in[1], out[0]); });
    GUARD_END(ctx)
}

int pgpu_encrypt_with_r_sk(pgpu_ctx* ctx, size_t count, const void* m, const void* r, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && r && c)), "pgpu_encrypt_with_r_sk: null argument");
    if (!ctx->has_secret) return fail(ctx, PGPU_ERR_STATE, "EncryptWithR (secret key): no secret key loaded");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    const size_t wn = ctx->wn * 4, w2 = (size_t)ctx->m_n2.sh.S * 4;
    ChunkedIo io;
    io.add_in(m, wn); io.add_in(r, wn); io.add_out(c, w2);
    io.align = resident_groups(ctx, ctx->m_p2.ready ? ctx->m_p2 : ctx->m_n2);
    return run_chunked(ctx, count, io, [&](size_t n, const uint32_t* const* in, uint32_t* const* out) { return encrypt_crt_dev(ctx, n, in[0], in[1], out[0]); });
    GUARD_END(ctx)
}

int pgpu_encrypt_with_rn(pgpu_ctx* ctx, size_t count, const void* m, const void* rn, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && rn && c)), "pgpu_encrypt_with_rn: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    const size_t wn = ctx->wn * 4, w2 = (size_t)ctx->m_n2.sh.S * 4;
    ChunkedIo io;
    io.add_in(m, wn); io.add_in(rn, w2); io.add_out(c, w2);
    io.align = resident_groups(ctx, ctx->m_n2);
    return run_chunked(ctx, count, io, [&](size_t n, const uint32_t* const* in, uint32_t* const* out) { return encrypt_rn_dev(ctx, n, in[0], in[1], out[0]); });
    GUARD_END(ctx)
}

int pgpu_decrypt(pgpu_ctx* ctx, size_t count, const void* c, void* m) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && c)), "pgpu_decrypt: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    const size_t wn = ctx->wn * 4, w2 = (size_t)ctx->m_n2.sh.S * 4;
    ChunkedIo io;
    io.add_in(c, w2); io.add_out(m, wn);
    io.align = resident_groups(ctx, ctx->m_p2.ready ? ctx->m_p2 : ctx->m_n2);
    return run_chunked(ctx, count, io, [&](size_t n, const uint32_t* const* in, uint32_t* const* out) { return decrypt_dev(ctx, n, in[0], out[0]); });
    GUARD_END(ctx)
}

int pgpu_partial_decrypt(pgpu_ctx* ctx, size_t count, const void* c, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (out && c)), "pgpu_partial_decrypt: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    const size_t w2 = (size_t)ctx->m_n2.sh.S * 4;
    ChunkedIo io;
    io.add_in(c, w2); io.add_out(out, w2);
    io.align = resident_groups(ctx, ctx->m_n2);
    return run_chunked(ctx, count, io, [&](size_t n, const uint32_t* const* in, uint32_t* const* o) { return pdec_dev(ctx, n, in[0], o[0]); });
    GUARD_END(ctx)
}

int pgpu_modexp(pgpu_ctx* ctx, int modsel, size_t count, const void* base, const void* exp, size_t exp_bytes, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (base && exp && out)), "pgpu_modexp: null argument");
    REQUIRE(ctx, exp_bytes > 0 && exp_bytes % 4 == 0, "pgpu_modexp: exp_bytes must be a positive multiple of 4");
    ModCtx* M = select_mod(ctx, modsel);
    if (!M) return fail(ctx, PGPU_ERR_UNSUPPORTED, "pgpu_modexp: modulus not available for this key");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    const size_t w = (size_t)M->sh.S * 4;
    ChunkedIo io;
    io.add_in(base, w); io.add_in(exp, exp_bytes); io.add_out(out, w);
    io.align = resident_groups(ctx, *M);
    return run_chunked(ctx, count, io, [&](size_t n, const uint32_t* const* in, uint32_t* const* o) {
        return modexp_items_dev(ctx, *M, n, in[0], in[1], (uint32_t)(exp_bytes / 4), o[0]); });
    GUARD_END(ctx)
}

int pgpu_const_mult(pgpu_ctx* ctx, size_t count, const void* c, const void* k, size_t k_bytes, void* out) {
    return pgpu_modexp(ctx, PGPU_MOD_N2, count, c, k, k_bytes, out);
}

int pgpu_modexp_shared(pgpu_ctx* ctx, int modsel, size_t count, const void* base, const uint8_t* exp_be, size_t exp_len, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (base && out)) && (exp_be || exp_len == 0), "pgpu_modexp_shared: null argument");
    ModCtx* M = select_mod(ctx, modsel);
    if (!M) return fail(ctx, PGPU_ERR_UNSUPPORTED, "pgpu_modexp_shared: modulus not available for this key");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    const BigU e = BigU::from_be(exp_be, exp_len);
    const size_t w = (size_t)M->sh.S * 4;
    ChunkedIo io;
    io.add_in(base, w); io.add_out(out, w);
    io.align = resident_groups(ctx, *M);
    return run_chunked(ctx, count, io, [&](size_t n, const uint32_t* const* in, uint32_t* const* o) { return modexp_shared_dev(ctx, *M, n, in[0], e, o[0]); });
    GUARD_END(ctx)
}

int pgpu_modmul(pgpu_ctx* ctx, int modsel, size_t count, const void* a, const void* b, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (a && b && out)), "pgpu_modmul: null argument");
    ModCtx* M = select_mod(ctx, modsel);
    if (!M) return fail(ctx, PGPU_ERR_UNSUPPORTED, "pgpu_modmul: modulus not available for this key");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    const size_t w = (size_t)M->sh.S * 4;
    ChunkedIo io;
    io.add_in(a, w); io.add_in(b, w); io.add_out(out, w);
    io.align = resident_groups(ctx, *M);
    return run_chunked(ctx, count, io, [&](size_t n, const uint32_t* const* in, uint32_t* const* o) { return modmul_dev(ctx, *M, n, in[0], in[1], o[0]); });
    GUARD_END(ctx)
}

int pgpu_add_pairs(pgpu_ctx* ctx, size_t count, const void* a, const void* b, void* out) {
    return pgpu_modmul(ctx, PGPU_MOD_N2, count, a, b, out);
}

int pgpu_add_reduce(pgpu_ctx* ctx, size_t count, const void* c, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && out && (count == 0 || c), "pgpu_add_reduce: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t w2 = (size_t)ctx->m_n2.sh.S * 4;
    uint32_t* dc = io.in(0, c ? c : out, std::max<size_t>(count, 1) * w2);
    uint32_t* dout = io.out(1, w2);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = prod_dev(ctx, ctx->m_n2, count, dc, dout))) return rc; }
    return io.finish(out, dout, w2);
    GUARD_END(ctx)
}

int pgpu_add_reduce_at_level(pgpu_ctx* ctx, int level, size_t count, const void* c, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && out && (count == 0 || c), "pgpu_add_reduce_at_level: null argument");
    REQUIRE(ctx, level == 1 || level == 2, "pgpu_add_reduce_at_level: level must be 1 or 2");
    ModCtx* M = select_mod(ctx, level == 1 ? PGPU_MOD_N2 : PGPU_MOD_N3);
    if (!M) return fail(ctx, PGPU_ERR_UNSUPPORTED, "pgpu_add_reduce_at_level: modulus not available for this key");
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t w = (size_t)M->sh.S * 4;
    uint32_t* dc = io.in(0, c ? c : out, std::max<size_t>(count, 1) * w);
    uint32_t* dout = io.out(1, w);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = prod_dev(ctx, *M, count, dc, dout))) return rc; }
    return io.finish(out, dout, w);
    GUARD_END(ctx)
}

int pgpu_dot_u64(pgpu_ctx* ctx, size_t count, const void* c, const uint64_t* k, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && out && (count == 0 || (c && k)), "pgpu_dot_u64: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t w2 = (size_t)ctx->m_n2.sh.S * 4;
    uint32_t* dc = io.in(0, c ? c : out, std::max<size_t>(count, 1) * w2);
    uint32_t* dk = io.in(1, k ? (const void*)k : (const void*)out, std::max<size_t>(count, 1) * 8);
    uint32_t* dout = io.out(5, w2);
    if (io.rc) return io.rc;
    if ((rc = pgpu_dot_u64_dev(ctx, count, dc, (const uint64_t*)dk, dout))) return rc;
    return io.finish(out, dout, w2);
    GUARD_END(ctx)
}

int pgpu_ctx_z_width(const pgpu_ctx* ctx, size_t* w_z) {
    if (!ctx || !w_z) return fail(nullptr, PGPU_ERR_ARG, "null argument");
    *w_z = (size_t)z_limbs(ctx) * 4;
    return PGPU_OK;
}

int pgpu_modinv(pgpu_ctx* ctx, int modsel, size_t count, const void* a, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (a && out)), "pgpu_modinv: null argument");
    ModCtx* M = select_mod(ctx, modsel);
    if (!M) return fail(ctx, PGPU_ERR_UNSUPPORTED, "pgpu_modinv: modulus not available for this key");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t w = (size_t)M->sh.S * 4;
    uint32_t* da = io.in(0, a, count * w);
    uint32_t* dout = io.out(1, count * w);
    uint32_t* dbad = io.out(2, 4);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = modinv_batch_dev(ctx, *M, count, da, dout, dbad))) return rc; }
    uint32_t bad = 0;
    if ((rc = io.finish(&bad, dbad, 4))) return rc;
    if ((rc = io.finish(out, dout, count * w))) return rc;
    if (bad != 0xffffffffu) return fail(ctx, PGPU_ERR_NOT_INVERTIBLE, "ModInverse: item " + std::to_string(bad) + " is not invertible");
    return PGPU_OK;
    GUARD_END(ctx)
}

int pgpu_sub_pairs(pgpu_ctx* ctx, size_t count, const void* a, const void* b, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (a && b && out)), "pgpu_sub_pairs: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    ModCtx& M = ctx->m_n2;
    HostIo io(ctx);
    const size_t w = (size_t)M.sh.S * 4;
    uint32_t* da = io.in(0, a, count * w);
    uint32_t* db = io.in(1, b, count * w);
    uint32_t* dout = io.out(2, count * w);
    uint32_t* dbad = io.out(3, 4);
    if (io.rc) return io.rc;
    {
        TimedScope ts(ctx);
        if ((rc = modinv_batch_dev(ctx, M, count, db, dout, dbad))) return rc;          // neg := ModInverse(c.C, ns1)  operations.go:43
        if ((rc = modmul_dev(ctx, M, count, da, dout, dout))) return rc;          // Mod(Mul(accumulator, neg))   :44-47
    }
    uint32_t bad = 0;
    if ((rc = io.finish(&bad, dbad, 4))) return rc;
    if ((rc = io.finish(out, dout, count * w))) return rc;
    if (bad != 0xffffffffu) return fail(ctx, PGPU_ERR_NOT_INVERTIBLE, "Sub: ciphertext " + std::to_string(bad) + " is not invertible mod n^2");
    return PGPU_OK;
    GUARD_END(ctx)
}

int pgpu_pdec_zkp_prove(pgpu_ctx* ctx, size_t count, const void* c, const void* r, void* dec, void* e, void* z) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && r && dec && e && z)), "pgpu_pdec_zkp_prove: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t w2 = (size_t)ctx->m_n2.sh.S * 4, wz = (size_t)z_limbs(ctx) * 4;
    uint32_t* dc = io.in(0, c, count * w2);
    uint32_t* dr = io.in(1, r, count * w2);
    uint32_t* ddec = io.out(2, count * w2);
    uint32_t* de = io.out(3, count * 32);
    uint32_t* dz = io.out(4, count * wz);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = zkp_prove_dev(ctx, count, dc, dr, ddec, de, dz))) return rc; }
    if ((rc = io.finish(dec, ddec, count * w2))) return rc;
    if ((rc = io.finish(e, de, count * 32))) return rc;
    return io.finish(z, dz, count * wz);
    GUARD_END(ctx)
}

int pgpu_pdec_zkp_verify(pgpu_ctx* ctx, size_t count, int id, const void* c, const void* dec, const void* e, const void* z, uint8_t* ok) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && dec && e && z && ok)), "pgpu_pdec_zkp_verify: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t w2 = (size_t)ctx->m_n2.sh.S * 4, wz = (size_t)z_limbs(ctx) * 4;
    uint32_t* dc = io.in(0, c, count * w2);
    uint32_t* ddec = io.in(1, dec, count * w2);
    uint32_t* de = io.in(2, e, count * 32);
    uint32_t* dz = io.in(3, z, count * wz);
    uint32_t* dok = io.out(4, count);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = zkp_verify_dev(ctx, count, id, dc, ddec, de, dz, (uint8_t*)dok))) return rc; }
    return io.finish(ok, dok, count);
    GUARD_END(ctx)
}

int pgpu_combine(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, void* m) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && k >= 0 && (k == 0 || ids) && (count == 0 || k == 0 || (decs && m)), "pgpu_combine: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t w2 = (size_t)ctx->m_n2.sh.S * 4, wn = ctx->wn * 4;
    uint32_t* dd = io.in(0, decs ? decs : (const void*)ids, std::max<size_t>(count * (size_t)k, 1) * (decs ? w2 : 1));
    uint32_t* dm = io.out(1, std::max<size_t>(count, 1) * wn);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = combine_dev(ctx, count, k, ids, dd, dm))) return rc; }
    if (count == 0) return PGPU_OK;
    return io.finish(m, dm, count * wn);
    GUARD_END(ctx)
}

int pgpu_pdec_zkp_prove_dev(pgpu_ctx* ctx, size_t count, const void* c, const void* r, void* dec, void* e, void* z) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && r && dec && e && z)), "pgpu_pdec_zkp_prove_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return zkp_prove_dev(ctx, count, (const uint32_t*)c, (const uint32_t*)r, (uint32_t*)dec, (uint32_t*)e, (uint32_t*)z);
    GUARD_END(ctx)
}

int pgpu_pdec_zkp_prove_given_dev(pgpu_ctx* ctx, size_t count, const void* c, const void* r, const void* dec, void* e, void* z) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && r && dec && e && z)), "pgpu_pdec_zkp_prove_given_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return zkp_prove_dev(ctx, count, (const uint32_t*)c, (const uint32_t*)r, (uint32_t*)dec, (uint32_t*)e, (uint32_t*)z, true);
    GUARD_END(ctx)
}

int pgpu_combine_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, void* m) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && k >= 0 && (k == 0 || ids), "pgpu_combine_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return combine_dev(ctx, count, k, ids, (const uint32_t*)decs, (uint32_t*)m);
    GUARD_END(ctx)
}


// ---- level 2, alternative encryption, randomness, nested operations, DDLEQ (protocols.cu)
int pgpu_ctx_set_alt_generator(pgpu_ctx* ctx, const uint8_t* h_be, size_t h_len, unsigned k_bits) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && h_be && h_len > 0, "pgpu_ctx_set_alt_generator: null argument");
    REQUIRE(ctx, k_bits >= 1 && k_bits <= ctx->wn * 32, "pgpu_ctx_set_alt_generator: K = 2^k_bits must fit an n-width record");
    int rc; if ((rc = set_device(ctx))) return rc;
    ctx->alt_h = BigU::from_be(h_be, h_len);
    REQUIRE(ctx, !ctx->alt_h.is_zero() && ctx->alt_h < ctx->n, "pgpu_ctx_set_alt_generator: H must be in [1, n)");
    ctx->alt_kbits = k_bits;
    return setup_alt(ctx);
    GUARD_END(ctx)
}

static int level_widths(pgpu_ctx* ctx, int level, size_t* w_m, size_t* w_c) {
    const size_t wn = ctx->wn * 4;
    if (level == 1) { *w_m = wn; *w_c = (size_t)ctx->m_n2.sh.S * 4; return PGPU_OK; }
    if (level == 2) {
        if (!ctx->level2_ready) return fail(ctx, PGPU_ERR_UNSUPPORTED, "level 2: n^3 is wider than the built kernel shapes");
        *w_m = 2 * wn; *w_c = (size_t)ctx->m_n3.sh.S * 4; return PGPU_OK;
    }
    return fail(ctx, PGPU_ERR_ARG, "encryption level must be 1 (EncLevelOne) or 2 (EncLevelTwo)");
}

int pgpu_encrypt_with_r_at_level(pgpu_ctx* ctx, int level, size_t count, const void* m, const void* r, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && r && c)), "pgpu_encrypt_with_r_at_level: null argument");
    if (level == 1) return pgpu_encrypt_with_r(ctx, count, m, r, c);
    size_t wm, wc; int rc;
    if ((rc = level_widths(ctx, level, &wm, &wc))) return rc;
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    uint32_t* dm = io.in(0, m, count * wm);
    uint32_t* dr = io.in(1, r, count * ctx->wn * 4);
    uint32_t* dc = io.out(2, count * wc);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = encrypt2_dev(ctx, count, dm, dr, dc))) return rc; }
    return io.finish(c, dc, count * wc);
    GUARD_END(ctx)
}

int pgpu_encrypt_with_r_at_level_sk(pgpu_ctx* ctx, int level, size_t count, const void* m, const void* r, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && r && c)), "pgpu_encrypt_with_r_at_level_sk: null argument");
    if (level == 1) return pgpu_encrypt_with_r_sk(ctx, count, m, r, c);
    size_t wm, wc; int rc;
    if ((rc = level_widths(ctx, level, &wm, &wc))) return rc;
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    uint32_t* dm = io.in(0, m, count * wm);
    uint32_t* dr = io.in(1, r, count * ctx->wn * 4);
    uint32_t* dc = io.out(2, count * wc);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = encrypt2_crt_dev(ctx, count, dm, dr, dc))) return rc; }
    return io.finish(c, dc, count * wc);
    GUARD_END(ctx)
}

int pgpu_alt_encrypt_with_r_at_level(pgpu_ctx* ctx, int level, size_t count, const void* m, const void* r, void* c) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && r && c)), "pgpu_alt_encrypt_with_r_at_level: null argument");
    size_t wm, wc; int rc;
    if ((rc = level_widths(ctx, level, &wm, &wc))) return rc;
    if (!ctx->has_alt) return fail(ctx, PGPU_ERR_STATE, "AltEncrypt: H and K were not loaded (pgpu_ctx_set_alt_generator)");
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    uint32_t* dm = io.in(0, m, count * wm);
    uint32_t* dr = io.in(1, r, count * ctx->wn * 4);
    uint32_t* dc = io.out(2, count * wc);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = alt_encrypt_dev(ctx, level, count, dm, dr, dc))) return rc; }
    return io.finish(c, dc, count * wc);
    GUARD_END(ctx)
}

int pgpu_decrypt_at_level(pgpu_ctx* ctx, int level, size_t count, const void* c, void* m) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (m && c)), "pgpu_decrypt_at_level: null argument");
    if (level == 1) return pgpu_decrypt(ctx, count, c, m);
    size_t wm, wc; int rc;
    if ((rc = level_widths(ctx, level, &wm, &wc))) return rc;
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    uint32_t* dc = io.in(0, c, count * wc);
    uint32_t* dm = io.out(1, count * wm);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = decrypt2_dev(ctx, count, dc, dm))) return rc; }
    return io.finish(m, dm, count * wm);
    GUARD_END(ctx)
}

int pgpu_randomize_with_r(pgpu_ctx* ctx, size_t count, const void* c, const void* r, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && r && out)), "pgpu_randomize_with_r: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t wn = ctx->wn * 4, w2 = (size_t)ctx->m_n2.sh.S * 4;
    uint32_t* dc = io.in(0, c, count * w2);
    uint32_t* dr = io.in(1, r, count * wn);
    uint32_t* dout = io.out(2, count * w2);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = randomize_dev(ctx, count, dc, dr, dout))) return rc; }
    return io.finish(out, dout, count * w2);
    GUARD_END(ctx)
}

int pgpu_extract_randomness(pgpu_ctx* ctx, int level, size_t count, const void* c, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && out)), "pgpu_extract_randomness: null argument");
    size_t wm, wc; int rc;
    if ((rc = level_widths(ctx, level, &wm, &wc))) return rc;
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    uint32_t* dc = io.in(0, c, count * wc);
    uint32_t* dout = io.out(1, count * ctx->wn * 4);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = extract_randomness_dev(ctx, level, count, dc, dout))) return rc; }
    return io.finish(out, dout, count * ctx->wn * 4);
    GUARD_END(ctx)
}

int pgpu_nested_randomize_with(pgpu_ctx* ctx, size_t count, const void* ct, const void* a, const void* b, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (ct && a && b && out)), "pgpu_nested_randomize_with: null argument");
    size_t wm, wc; int rc;
    if ((rc = level_widths(ctx, 2, &wm, &wc))) return rc;
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t wn = ctx->wn * 4;
    uint32_t* dct = io.in(0, ct, count * wc);
    uint32_t* da = io.in(1, a, count * wn);
    uint32_t* db = io.in(2, b, count * wn);
    uint32_t* dout = io.out(3, count * wc);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = nested_randomize_dev(ctx, count, dct, da, db, dout))) return rc; }
    return io.finish(out, dout, count * wc);
    GUARD_END(ctx)
}

static int nested_addsub(pgpu_ctx* ctx, bool sub, size_t count, const void* ct1, const void* ct2, void* out) {
    size_t wm, wc; int rc;
    if ((rc = level_widths(ctx, 2, &wm, &wc))) return rc;
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const uint32_t S2 = ctx->m_n2.sh.S, S3 = ctx->m_n3.sh.S;
    uint32_t* d1 = io.in(0, ct1, count * wc);
    uint32_t* d2 = io.in(1, ct2, count * (size_t)S2 * 4);
    uint32_t* dout = io.out(2, count * wc);
    uint32_t* dinv = io.out(3, count * (size_t)S2 * 4);
    uint32_t* dbad = io.out(4, 4);
    if (io.rc) return io.rc;
    const uint32_t* e = d2;
    {
        TimedScope ts(ctx);
        if (sub) { if ((rc = modinv_batch_dev(ctx, ctx->m_n2, count, d2, dinv, dbad))) return rc; e = dinv; }        // operations.go:137
        if ((rc = modexp_items_io(ctx, ctx->m_n3, count, IoDesc{d1, S3, S3}, ExpDesc{e, S2, (uint32_t)ctx->n2.bitlen(), nullptr}, dout))) return rc;   // ConstMult :126,139
    }
    uint32_t bad = 0xffffffffu;
    if (sub && (rc = io.finish(&bad, dbad, 4))) return rc;
    if ((rc = io.finish(out, dout, count * wc))) return rc;
    if (bad != 0xffffffffu) return fail(ctx, PGPU_ERR_NOT_INVERTIBLE, "NestedSub: ciphertext " + std::to_string(bad) + " is not invertible mod n^2");
    return PGPU_OK;
}

int pgpu_nested_add(pgpu_ctx* ctx, size_t count, const void* ct1, const void* ct2, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (ct1 && ct2 && out)), "pgpu_nested_add: null argument");
    return nested_addsub(ctx, false, count, ct1, ct2, out);
    GUARD_END(ctx)
}

int pgpu_nested_sub(pgpu_ctx* ctx, size_t count, const void* ct1, const void* ct2, void* out) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (ct1 && ct2 && out)), "pgpu_nested_sub: null argument");
    return nested_addsub(ctx, true, count, ct1, ct2, out);
    GUARD_END(ctx)
}

int pgpu_ddleq_prove(pgpu_ctx* ctx, size_t count, unsigned secpar, const void* ct1, const void* ct2, const void* a, const void* b,
                     const void* x, const void* y, void* alpha, void* e, void* f) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (ct1 && ct2 && a && b && x && y && alpha && e && f)), "pgpu_ddleq_prove: null argument");
    REQUIRE(ctx, secpar >= 1, "pgpu_ddleq_prove: secpar must be positive");
    size_t wm, w3; int rc;
    if ((rc = level_widths(ctx, 2, &wm, &w3))) return rc;
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t wn = ctx->wn * 4, w2 = 2 * wn, total = count * secpar;
    uint32_t* d1 = io.in(0, ct1, count * w3);
    uint32_t* d2 = io.in(1, ct2, count * w3);
    uint32_t* da = io.in(2, a, count * wn);
    uint32_t* db = io.in(3, b, count * wn);
    uint32_t* dx = io.in(4, x, total * wn);
    uint32_t* dy = io.in(5, y, total * wn);
    uint32_t* dal = io.out(6, total * w3);
    uint32_t* de = io.out(7, total * w2);
    uint32_t* df = io.out(8, total * w3);
    uint32_t* dbad = io.out(9, 4);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = ddleq_prove_dev(ctx, count, secpar, d1, d2, da, db, dx, dy, dal, de, df, dbad))) return rc; }
    uint32_t bad = 0;
    if ((rc = io.finish(&bad, dbad, 4))) return rc;
    if (bad != 0xffffffffu)     // the reference panics here (ddleq.go:67-69)
        return fail(ctx, PGPU_ERR_ARG, "cannot prove re-encryption because inputs are wrong (statement " + std::to_string(bad) + ")");
    if ((rc = io.finish(alpha, dal, total * w3))) return rc;
    if ((rc = io.finish(e, de, total * w2))) return rc;
    return io.finish(f, df, total * w3);
    GUARD_END(ctx)
}

int pgpu_ddleq_verify(pgpu_ctx* ctx, size_t count, unsigned secpar, const void* ct1, const void* ct2, const void* x, const void* y,
                      const void* alpha, const void* e, const void* f, uint8_t* ok) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (ct1 && ct2 && x && y && alpha && e && f && ok)), "pgpu_ddleq_verify: null argument");
    REQUIRE(ctx, secpar >= 1, "pgpu_ddleq_verify: secpar must be positive");
    size_t wm, w3; int rc;
    if ((rc = level_widths(ctx, 2, &wm, &w3))) return rc;
    if (count == 0) return PGPU_OK;
    if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t wn = ctx->wn * 4, w2 = 2 * wn, total = count * secpar;
    uint32_t* d1 = io.in(0, ct1, count * w3);
    uint32_t* d2 = io.in(1, ct2, count * w3);
    uint32_t* dx = io.in(2, x, total * wn);
    uint32_t* dy = io.in(3, y, total * wn);
    uint32_t* dal = io.in(4, alpha, total * w3);
    uint32_t* de = io.in(5, e, total * w2);
    uint32_t* df = io.in(6, f, total * w3);
    uint32_t* dok = io.out(7, total);
    if (io.rc) return io.rc;
    { TimedScope ts(ctx); if ((rc = ddleq_verify_dev(ctx, count, secpar, d1, d2, dx, dy, dal, de, df, (uint8_t*)dok))) return rc; }
    return io.finish(ok, dok, total);
    GUARD_END(ctx)
}

int pgpu_combine_strided_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, size_t share_stride, void* m) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && k >= 0 && (k == 0 || ids), "pgpu_combine_strided_dev: null argument");
    REQUIRE(ctx, share_stride >= count, "pgpu_combine_strided_dev: share_stride must be at least count");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return combine_dev(ctx, count, k, ids, (const uint32_t*)decs, (uint32_t*)m, share_stride);
    GUARD_END(ctx)
}

int pgpu_pdec_zkp_verify_shared_dev(pgpu_ctx* ctx, size_t n, int k, const int* ids, const void* c, const void* dec, const void* e, const void* z,
                                    uint8_t* ok) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && k >= 1 && ids && (n == 0 || (c && dec && e && z && ok)), "pgpu_pdec_zkp_verify_shared_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return zkp_verify_shared_dev(ctx, n, k, ids, (const uint32_t*)c, (const uint32_t*)dec, (const uint32_t*)e, (const uint32_t*)z, ok);
    GUARD_END(ctx)
}

int pgpu_combine_verified_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, size_t share_stride, const uint8_t* ok,
                              void* m, uint8_t* item_ok) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && k >= 0 && (k == 0 || ids) && (count == 0 || (decs && ok && m)), "pgpu_combine_verified_dev: null argument");
    REQUIRE(ctx, share_stride == 0 || share_stride >= count, "pgpu_combine_verified_dev: share_stride must be at least count");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    size_t failed = 0;
    if ((rc = combine_verified_dev(ctx, count, k, ids, (const uint32_t*)decs, share_stride, ok, (uint32_t*)m, item_ok, &failed))) return rc;
    if (failed) return fail(ctx, PGPU_ERR_THRESHOLD, "Threshold not meet for " + std::to_string(failed) + " of " + std::to_string(count) + " ciphertexts");
    return PGPU_OK;
    GUARD_END(ctx)
}

int pgpu_combine_verified(pgpu_ctx* ctx, size_t count, int k, const int* ids, const void* decs, const uint8_t* ok, void* m, uint8_t* item_ok) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && k >= 0 && (k == 0 || ids) && (count == 0 || k == 0 || (decs && ok)) && (count == 0 || m), "pgpu_combine_verified: null argument");
    if (count == 0) return PGPU_OK;
    int rc; if ((rc = set_device(ctx))) return rc;
    HostIo io(ctx);
    const size_t w2 = (size_t)ctx->m_n2.sh.S * 4, wn = ctx->wn * 4;
    uint32_t* dd = io.in(0, k ? decs : (const void*)m, std::max<size_t>(count * (size_t)k, 1) * (k ? w2 : 1));
    uint32_t* dok = io.in(1, k ? (const void*)ok : (const void*)m, std::max<size_t>(count * (size_t)k, 1));
    uint32_t* dm = io.out(2, count * wn);
    uint32_t* dit = io.out(3, count);
    if (io.rc) return io.rc;
    size_t failed = 0;
    { TimedScope ts(ctx); if ((rc = combine_verified_dev(ctx, count, k, ids, dd, count, (const uint8_t*)dok, dm, (uint8_t*)dit, &failed))) return rc; }
    if (item_ok) CU(ctx, cudaMemcpyAsync(item_ok, dit, count, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = io.finish(m, dm, count * wn))) return rc;
    if (failed) return fail(ctx, PGPU_ERR_THRESHOLD, "Threshold not meet for " + std::to_string(failed) + " of " + std::to_string(count) + " ciphertexts");
    return PGPU_OK;
    GUARD_END(ctx)
}

int pgpu_pdec_zkp_verify_dev(pgpu_ctx* ctx, size_t count, int id, const void* c, const void* dec, const void* e, const void* z, uint8_t* ok) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && (count == 0 || (c && dec && e && z && ok)), "pgpu_pdec_zkp_verify_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return zkp_verify_dev(ctx, count, id, (const uint32_t*)c, (const uint32_t*)dec, (const uint32_t*)e, (const uint32_t*)z, ok);
    GUARD_END(ctx)
}

int pgpu_pdec_zkp_verify_multi_dev(pgpu_ctx* ctx, size_t n_per_id, int k, const int* ids, const void* c, const void* dec, const void* e,
                                   const void* z, uint8_t* ok) {
    GUARD_BEGIN
    REQUIRE(ctx, ctx && k >= 0 && (k == 0 || ids) && (n_per_id == 0 || k == 0 || (c && dec && e && z && ok)), "pgpu_pdec_zkp_verify_multi_dev: null argument");
    int rc; if ((rc = set_device(ctx))) return rc;
    TimedScope ts(ctx);
    return zkp_verify_multi_dev(ctx, n_per_id, k, ids, (const uint32_t*)c, (const uint32_t*)dec, (const uint32_t*)e, (const uint32_t*)z, ok);
    GUARD_END(ctx)
}

int pgpu_ctx_launch_count(const pgpu_ctx* ctx, uint64_t* launches) {
    if (!ctx || !launches) return fail(nullptr, PGPU_ERR_ARG, "null argument");
    *launches = ctx->launches;
    return PGPU_OK;
}

int pgpu_ctx_program_cost(const pgpu_ctx* ctx, int what, uint32_t* limbs, uint32_t* n_sqr, uint32_t* n_mul) {
    if (!ctx) return fail(nullptr, PGPU_ERR_ARG, "null context");
    uint32_t S = 0, sq = 0, mu = 0;
    switch (what) {
        case 0: S = ctx->m_n2.sh.S; sq = ctx->prog_enc.n_sqr; mu = ctx->prog_enc.n_mul; break;
        case 1:
            if (!ctx->has_secret) return fail(nullptr, PGPU_ERR_STATE, "no secret key");
            S = ctx->m_p2.sh.S; sq = ctx->prog_dec_p.n_sqr + ctx->prog_dec_q.n_sqr; mu = ctx->prog_dec_p.n_mul + ctx->prog_dec_q.n_mul; break;
        case 2:
            if (!ctx->has_share) return fail(nullptr, PGPU_ERR_STATE, "no share");
            S = ctx->m_n2.sh.S; sq = ctx->prog_pdec.n_sqr; mu = ctx->prog_pdec.n_mul; break;
        case 3:
            if (!ctx->has_enc_crt) return fail(nullptr, PGPU_ERR_STATE, "no secret-key encryption programs");
            S = ctx->m_p2.sh.S; sq = ctx->prog_encq.n_sqr + ctx->prog_encp.n_sqr; mu = ctx->prog_encq.n_mul + ctx->prog_encp.n_mul; break;
        default: return fail(nullptr, PGPU_ERR_ARG, "bad selector");
    }
    if (limbs) *limbs = S;
    if (n_sqr) *n_sqr = sq;
    if (n_mul) *n_mul = mu;
    return PGPU_OK;
}

int pgpu_ctx_kernel_shape(const pgpu_ctx* ctx, int modsel, int* tpi, int* limbs_per_lane, int* fp64, int* resident_groups) {
    if (!ctx) return fail(nullptr, PGPU_ERR_ARG, "null context");
    const ModCtx* m = nullptr;
    switch (modsel) {
        case 0: m = &ctx->m_n; break;
        case 1: m = &ctx->m_n2; break;
        case 2: m = &ctx->m_n3; break;
        case 3: m = &ctx->m_p2; break;
        case 4: m = &ctx->m_p3; break;
        default: return fail(nullptr, PGPU_ERR_ARG, "bad modulus selector");
    }
    if (!m->ready) return fail(nullptr, PGPU_ERR_STATE, "modulus not loaded");
    if (tpi) *tpi = m->sh.tpi;
    if (limbs_per_lane) *limbs_per_lane = m->sh.L;
    if (fp64) *fp64 = m->sh.fp64 ? 1 : 0;
    if (resident_groups) *resident_groups = ctx->sms * m->blocks_per_sm * (VM_BLOCK_THREADS / m->sh.tpi);
    return PGPU_OK;
}

int pgpu_ctx_enable_timing(pgpu_ctx* ctx, int on) {
    if (!ctx) return fail(nullptr, PGPU_ERR_ARG, "null context");
    ctx->timing = on != 0; ctx->ev_valid = false;
    return PGPU_OK;
}

int pgpu_ctx_last_kernel_ms(pgpu_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(nullptr, PGPU_ERR_ARG, "null argument");
    if (!ctx->ev_valid) return fail(ctx, PGPU_ERR_STATE, "no timed call recorded");
    CU(ctx, cudaEventSynchronize(ctx->ev1));
    CU(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return PGPU_OK;
}

int pgpu_selftest_bn(int op, const uint8_t* a, size_t a_len, const uint8_t* b, size_t b_len,
                     const uint8_t* m, size_t m_len, uint8_t* out, size_t* out_len) {
    GUARD_BEGIN
    REQUIRE(nullptr, out && out_len, "null argument");
    const BigU A = BigU::from_be(a, a_len), B = BigU::from_be(b, b_len), M = BigU::from_be(m, m_len);
    BigU r;
    switch (op) {
        case 0: r = A * B; break;
        case 1: r = A / B; break;
        case 2: r = A % B; break;
        case 3: if (!BigU::modinv(A, M, r)) return fail(nullptr, PGPU_ERR_NOT_INVERTIBLE, "not invertible"); break;
        case 4: r = BigU::modexp(A, B, M); break;
        case 5: r = BigU::isqrt(A); break;
        default: return fail(nullptr, PGPU_ERR_ARG, "bad op");
    }
    const size_t nbytes = (r.bitlen() + 7) / 8;
    if (nbytes > *out_len) return fail(nullptr, PGPU_ERR_ARG, "output buffer too small");
    for (size_t i = 0; i < nbytes; ++i) out[nbytes - 1 - i] = (uint8_t)(r.v[i / 4] >> (8 * (i % 4)));
    *out_len = nbytes;
    return PGPU_OK;
    GUARD_END(nullptr)
}

int pgpu_selftest_program(int kind, const uint8_t* mod_be, size_t mod_len, const uint8_t* base_be, size_t base_len,
                          const uint8_t* shared_exp_be, size_t shared_len, const uint32_t* item_exps, uint32_t exp_limbs, uint32_t k, uint32_t pre,
                          uint8_t* out, size_t out_cap, uint32_t* n_out, uint32_t* n_sqr, uint32_t* n_mul) {
    GUARD_BEGIN
    REQUIRE(nullptr, mod_be && base_be && out && n_out, "null argument");
    const BigU N = BigU::from_be(mod_be, mod_len), base = BigU::from_be(base_be, base_len), e = BigU::from_be(shared_exp_be, shared_len);
    REQUIRE(nullptr, !N.is_zero(), "zero modulus");
    Program P;
    switch (kind) {
        case 0: P.emit(OP_LDI, 0); P.emit(OP_MULC, K_R2); emit_pow_shared(P, e, 0); P.emit(OP_MULC, K_ONE); P.emit(OP_STO, 0); break;
        case 1: REQUIRE(nullptr, item_exps && exp_limbs, "no exponent"); P.emit(OP_LDI, 0); P.emit(OP_MULC, K_R2); emit_pow_items(P, 32 * exp_limbs, 0);
                P.emit(OP_MULC, K_ONE); P.emit(OP_STO, 0); break;
        case 2: REQUIRE(nullptr, item_exps && exp_limbs && k >= 1 && k <= 8, "1 to 8 exponents"); build_multi_program(P, k, 32 * exp_limbs, pre); break;
        case 3: REQUIRE(nullptr, item_exps && exp_limbs, "no exponent"); build_pdec_a_program(P, e, 32 * exp_limbs); break;
        default: return fail(nullptr, PGPU_ERR_ARG, "bad program kind");
    }
    P.ops.push_back(vm_op(OP_END, 0));
    selftest_last_ops() = P.ops;
    selftest_last_tbl() = P.tbl_entries;
    std::vector<BigU> outs;
    const int rc = host_vm_run(P, N, base, item_exps, exp_limbs, kind == 2 ? exp_limbs : 0u, outs);
    if (rc) return fail(nullptr, rc, "host interpreter: op not supported");
    REQUIRE(nullptr, outs.size() * mod_len <= out_cap, "output buffer too small");
    for (size_t j = 0; j < outs.size(); ++j) {
        std::vector<uint32_t> l = outs[j].limbs((mod_len + 3) / 4);
        for (size_t i = 0; i < mod_len; ++i) out[j * mod_len + (mod_len - 1 - i)] = (uint8_t)(l[i / 4] >> (8 * (i % 4)));
    }
    *n_out = (uint32_t)outs.size();
    if (n_sqr) *n_sqr = P.n_sqr;
    if (n_mul) *n_mul = P.n_mul;
    return PGPU_OK;
    GUARD_END(nullptr)
}

int pgpu_selftest_last_program(uint32_t* ops, size_t cap, size_t* n_ops, uint32_t* tbl_entries) {
    if (!n_ops) return fail(nullptr, PGPU_ERR_ARG, "null argument");
    const std::vector<uint32_t>& v = selftest_last_ops();
    *n_ops = v.size();
    if (tbl_entries) *tbl_entries = selftest_last_tbl();
    if (v.empty()) return fail(nullptr, PGPU_ERR_STATE, "no program compiled by pgpu_selftest_program on this thread yet");
    if (!ops || cap < v.size()) return fail(nullptr, PGPU_ERR_ARG, "ops buffer too small");
    std::copy(v.begin(), v.end(), ops);
    return PGPU_OK;
}

}  // extern "C"
#pragma GCC visibility pop
