// Internal interface of the engine behind include/pgpu.h: key contexts, the host-side compiler from
// reference functions to powm_vm micro-programs (vm.h) and the device-side operations the C ABI
// (capi.cu) and the protocol layers (protocols.cu) are assembled from.  Nothing here runs CPU
// arithmetic on a batch item: the host derives per-key constants and programs, then launches kernels.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <list>
#include <map>
#include <string>
#include <vector>

#include "../../include/pgpu.h"
#include "aux.h"
#include "bn_host.hpp"
#include "launch.h"
#include "vm.h"

namespace pgpu {

using Shape = VmShape;

// kconst slots shared by every modulus
enum : uint32_t {
    K_R2 = 0, K_R1 = 1, K_ONE = 2, K_NR2 = 4, K_FIX = 5,
    K_R3 = 3,     // W * R^2 mod N, W = 2^(32*S): second chunk of a record split at the record width -> Montgomery form
    K_NEG1 = 6,   // -1                         (Montgomery form; level-2 g^m shortcut)
    K_C1 = 7,     // n^2 / 2 mod n^3            (Montgomery form)
    K_NM = 8,     // n                          (Montgomery form)
    K_NSM = 9,    // n^s for the modulus n^(s+1) (Montgomery form; g^(-v) = g^(n^s - v))
    K_R4 = 10,    // W^2 * R^2 mod N: third chunk
    K_CRT = 11,   // secret-key EncryptWithR: q^-2 mod p^2 (plain) in the p^2 context, q^2 * R^2 mod n^2 in the n^2 context
    K_SLOTS = 16
};

struct Program {
    std::vector<uint32_t> ops;
    uint32_t* d_ops = nullptr;
    size_t d_bytes = 0;             // size of d_ops (scrubbed before it is freed: the ops of a secret exponent spell it out)
    uint32_t n_sqr = 0, n_mul = 0, tbl_entries = 0;
    bool per_item = false;          // uses OP_WIN: table entries picked by per-item exponent bits
    void emit(uint32_t code, uint32_t arg) { ops.push_back(vm_op(code, arg)); if (code == OP_WIN) per_item = true; }
    void use_slot(uint32_t s) { tbl_entries = std::max(tbl_entries, s + 1); }
};

struct ModCtx {
    bool ready = false;
    Shape sh{};
    BigU N, R1, R2, R3;             // R mod N, R^2 mod N, W*R^2 mod N with R the kernel's Montgomery radix, W = 2^(32*S)
    BigU W1;                        // W mod N: the radix of the 32-bit-limb helper kernels and of records split into chunks
    Shape sh32{};                   // integer-pipe shape of the same record width (prod_reduce_kernel)
    uint32_t np0 = 0;
    uint32_t* d_mod = nullptr;
    uint32_t* d_kconst = nullptr;   // K_SLOTS records of S limbs
    int blocks_per_sm = 0;
    Shape sh_items{};               // shape for programs with per-item exponents (more warps in flight hide the table loads)
    int blocks_per_sm_items = 0;
    uint64_t fix_T = 0;             // prod_dev: K_FIX currently holds W^(fix_T - 1) * R (0: not set)
};

// fixed-base comb table for OP_FIXW: row k holds base^(d * 2^(w*k)), d = 0 .. 2^w-1, Montgomery form
struct FixedTable {
    uint32_t* d = nullptr;
    uint32_t w = 0, nwin = 0, bits = 0;
};

}  // namespace pgpu

struct pgpu_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;

    pgpu::BigU n, n2, n3;
    size_t wn = 0;         // limbs of an n-width record
    pgpu::ModCtx m_n, m_n2, m_n3, m_p2, m_q2;

    // secret key
    bool has_secret = false;
    pgpu::BigU p, q;
    uint32_t* d_crt = nullptr; size_t d_crt_bytes = 0;
    uint32_t crt_np0_p = 0, crt_np0_q = 0;
    int crt_h = 0;

    // threshold key
    bool has_threshold = false, has_share = false;
    int tk_l = 0, tk_w = 0, tk_id = 0;
    pgpu::BigU tk_share, tk_v, tk_delta;
    std::vector<pgpu::BigU> tk_vi;

    pgpu::Program prog_enc, prog_dec_p, prog_dec_q, prog_pdec;
    pgpu::Program prog_encq, prog_encp, prog_encf;   // secret-key EncryptWithR over p^2, q^2 (encrypt_crt_dev)
    bool has_enc_crt = false;

    // level 2 (mod n^3), alternative encryption, randomness extraction (protocols.cu)
    bool level2_ready = false;
    pgpu::Program prog_enc2, prog_rand;
    uint32_t* d_rec2 = nullptr;             // constants of recover2_kernel
    pgpu::ModCtx m_p3, m_q3;                // CRT moduli of level-2 Decrypt
    pgpu::Program prog_dec2_p, prog_dec2_q;
    pgpu::Program prog_enc2q, prog_enc2p, prog_enc2f;   // secret-key EncryptWithRAtLevel(2) over p^3, q^3 (encrypt2_crt_dev)
    bool enc2_crt_ready = false;
    uint32_t* d_crt2 = nullptr; size_t d_crt2_bytes = 0, crt2_cq_off = 0, crt2_g_off = 0;
    uint32_t crt2_np0[4] = {0, 0, 0, 0};
    bool crt2_ready = false;
    bool has_alt = false;
    pgpu::BigU alt_h; uint32_t alt_kbits = 0;
    pgpu::FixedTable fix_h1, fix_h2, fix_v;
    pgpu::Program prog_alt1, prog_alt2;
    std::map<std::string, pgpu::Program> prog_cache;
    std::list<std::string> pows_lru;         // least recently used first: keys of the ad-hoc shared-exponent programs ("pows:...") in prog_cache

    // scratch
    uint32_t* d_table = nullptr; size_t table_limbs = 0;
    void* d_stage[16] = {};                 // 0-9: host-buffer staging (HostIo); 12-15: scratch of the device-side ops
    size_t stage_bytes[16] = {};

    // chunked host path (ChunkedIo): copy streams and the events that order chunk k's H2D, kernels and D2H
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {}, ev_cmp[2] = {}, ev_out[2] = {};

    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
};

// device buffer handle of the C ABI (pgpu_buf_alloc): owned by the context it was allocated on
struct pgpu_buf {
    pgpu_ctx* ctx = nullptr;
    void* ptr = nullptr;
    size_t bytes = 0;
};

namespace pgpu {

int fail(pgpu_ctx* ctx, int code, const std::string& msg);
std::string& thread_error();

#define CU(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return ::pgpu::fail(ctx, PGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#define REQUIRE(ctx, cond, msg) do { if (!(cond)) return ::pgpu::fail(ctx, PGPU_ERR_ARG, msg); } while (0)
#define GUARD_BEGIN try {
#define GUARD_END(ctx) } catch (const std::exception& ex) { return ::pgpu::fail(ctx, PGPU_ERR_ARG, ex.what()); }

struct IoDesc { const uint32_t* ptr; uint32_t stride, limbs; uint32_t div = 1; };   // item i reads record i / div
// per-item exponents (OP_WIN / OP_FIXW): records `stride` limbs apart of which the low `bits` bits count
struct ExpDesc { const uint32_t* ptr = nullptr; uint32_t stride = 0, bits = 0; const uint32_t* fixed = nullptr; uint32_t sub = 0; };   // sub: limbs between the exponents of one item (OP_BKT)

// stream-ordered temporary device buffer
struct DevBuf {
    pgpu_ctx* c; uint32_t* p = nullptr; cudaError_t err = cudaSuccess;
    DevBuf(pgpu_ctx* ctx, size_t limbs) : c(ctx) { err = cudaMallocAsync((void**)&p, std::max<size_t>(limbs, 1) * 4, ctx->stream); }
    ~DevBuf() { if (p) cudaFreeAsync(p, c->stream); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};
#define DEVBUF(name, ctx, limbs) ::pgpu::DevBuf name(ctx, limbs); if (name.err != cudaSuccess) return ::pgpu::fail(ctx, PGPU_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(name.err))

struct TimedScope {
    pgpu_ctx* c;
    explicit TimedScope(pgpu_ctx* ctx) : c(ctx) { if (c->timing) { cudaEventRecord(c->ev0, c->stream); } }
    ~TimedScope() { if (c->timing) { cudaEventRecord(c->ev1, c->stream); c->ev_valid = true; } }
};

// host-buffer staging for the blocking entry points
struct HostIo {
    pgpu_ctx* ctx;
    int rc = PGPU_OK;
    explicit HostIo(pgpu_ctx* c) : ctx(c) {}
    uint32_t* in(int slot, const void* host, size_t bytes);
    uint32_t* out(int slot, size_t bytes);
    int finish(void* host, const uint32_t* dev, size_t bytes);
};

// Chunked, double-buffered host path of the blocking entry points whose items are independent: chunk k's kernels run
// while chunk k+1 is copied in and chunk k-1 is copied out (three streams), and the staging memory is two chunks
// instead of the whole batch.  With pageable host memory the copies block the calling thread, so the loop issues the
// next H2D before the previous D2H (software pipelining) -- the kernels still overlap both.
struct ChunkedIo {
    static constexpr int MAX_IN = 3, MAX_OUT = 2;
    const void* in[MAX_IN] = {};   size_t in_w[MAX_IN] = {};   int n_in = 0;     // host pointers, bytes per item
    void* out[MAX_OUT] = {};       size_t out_w[MAX_OUT] = {}; int n_out = 0;
    size_t align = 1;              // chunk sizes are multiples of this (resident groups of the kernel's persistent grid)
    void add_in(const void* p, size_t w) { in[n_in] = p; in_w[n_in++] = w; }
    void add_out(void* p, size_t w) { out[n_out] = p; out_w[n_out++] = w; }
};
size_t chunk_items(size_t count, size_t align);
// fn(n, din, dout): enqueue the operation for n items on ctx->stream (device pointers of the chunk)
int run_chunked(pgpu_ctx* ctx, size_t count, const ChunkedIo& io,
                const std::function<int(size_t, const uint32_t* const*, uint32_t* const*)>& fn);
size_t resident_groups(const pgpu_ctx* ctx, const ModCtx& m);

bool pick_shape(size_t limbs, Shape& out);
int upload(pgpu_ctx* ctx, uint32_t* dst, const std::vector<uint32_t>& v);
// Device memory that held key material (CRT constants, Montgomery constants of p^2 / q^2, the window programs of secret
// exponents) or plaintexts (staging, scratch tables) is zeroed before it goes back to the allocator; the same for the
// context's long-lived host copies of p, q and the share.  Waits for the device first: a queued kernel may still read it.
void dev_scrub_free(void* p, size_t bytes);
void scrub(BigU& x);
void scrub(std::vector<uint32_t>& v);
int set_kconst(pgpu_ctx* ctx, ModCtx& m, uint32_t slot, const BigU& v);
int modctx_init(pgpu_ctx* ctx, ModCtx& m, const BigU& N);
void modctx_free(ModCtx& m);
int choose_window(size_t bits);
void emit_pow_shared(Program& P, const BigU& e, uint32_t tb);
void emit_pow_items(Program& P, size_t exp_bits, uint32_t tb);
void build_multi_program(Program& P, uint32_t k, uint32_t bits, uint32_t pre);
void build_pdec_a_program(Program& P, const BigU& e1, uint32_t rbits);
int program_upload(pgpu_ctx* ctx, Program& P);
void program_free(Program& P);
int ensure_table(pgpu_ctx* ctx, size_t limbs);
int run_vm(pgpu_ctx* ctx, const ModCtx& m, const Program& prog, size_t count,
           const IoDesc* ins, int n_in, uint32_t* out, uint32_t out_stride, uint32_t out_limbs,
           const ExpDesc& ex = ExpDesc(), uint32_t* out2 = nullptr, uint32_t out2_stride = 0, int force_blocks = 0);
int vm_full_blocks(const pgpu_ctx* ctx, const ModCtx& m);
int dot_pippenger_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, const uint32_t* k, uint32_t k_limbs, uint32_t* out);
int stage(pgpu_ctx* ctx, int slot, size_t bytes, void** out);
ModCtx* select_mod(pgpu_ctx* ctx, int modsel);
int build_encrypt(pgpu_ctx* ctx);
int build_pdec(pgpu_ctx* ctx);
int build_decrypt_half(pgpu_ctx* ctx, Program& P, const BigU& pm1);
int setup_crt(pgpu_ctx* ctx);
int encrypt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c);
int decrypt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, uint32_t* m);
int encrypt_crt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c);
int encrypt_rn_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* rn, uint32_t* c);
int pdec_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, uint32_t* out);
Program* cached_program(pgpu_ctx* ctx, const std::string& key);
int modexp_items_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* base, const uint32_t* exp, uint32_t exp_limbs, uint32_t* out,
                     bool broadcast_base = false);
int modexp_shared_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* base, const BigU& e, uint32_t* out);
// out[i*k + s] = (base[i]^pre)^(exp[i*k + s]) for k exponents per base (item-major records): the squarings are shared by the k
// exponentiations (right-to-left, bucket method), pre = 1, 2 or 4 squares the base first
int modexp_multi_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, uint32_t k, const uint32_t* base, uint32_t pre, const uint32_t* exp, uint32_t exp_limbs,
                     uint32_t* out);
int modmul_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* a, const uint32_t* b, uint32_t* out);
// the same three with explicit record descriptors (narrower records, broadcast, one record per `div` items)
int modexp_items_io(pgpu_ctx* ctx, const ModCtx& M, size_t count, const IoDesc& base, const ExpDesc& exp, uint32_t* out);
int modexp_shared_io(pgpu_ctx* ctx, const ModCtx& M, size_t count, const IoDesc& base, const BigU& e, uint32_t* out);
int modmul_io(pgpu_ctx* ctx, const ModCtx& M, size_t count, const IoDesc& a, const IoDesc& b, uint32_t* out);
int prod_dev(pgpu_ctx* ctx, ModCtx& M, size_t count, const uint32_t* in, uint32_t* out);
int modinv_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* in, uint32_t* out, uint32_t* d_first_bad);
int modinv_batch_dev(pgpu_ctx* ctx, const ModCtx& M, size_t count, const uint32_t* in, uint32_t* out, uint32_t* d_first_bad);
int bigmul_dev(pgpu_ctx* ctx, size_t count, const uint32_t* a, uint32_t na, const uint32_t* b, uint32_t nb, uint32_t* out);
int sha_dev(pgpu_ctx* ctx, size_t count, int n_seg, const uint32_t* const* seg, const uint32_t* stride, const int* limbs, uint32_t* out,
            const uint32_t* div = nullptr);
uint32_t z_limbs(const pgpu_ctx* ctx);
// dec_given: dec already holds PartialDecrypt(c) (the caller ran pdec_dev), only a, b, E, Z are computed
int zkp_prove_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, const uint32_t* r, uint32_t* dec, uint32_t* e, uint32_t* z, bool dec_given = false);
int zkp_verify_dev(pgpu_ctx* ctx, size_t count, int id, const uint32_t* c, const uint32_t* dec, const uint32_t* e, const uint32_t* z, uint8_t* ok);
// share j's batch starts at record j*share_stride (0 = count: tightly packed)
int zkp_verify_multi_dev(pgpu_ctx* ctx, size_t n_per_id, int k, const int* ids, const uint32_t* c, const uint32_t* dec, const uint32_t* e,
                         const uint32_t* z, uint8_t* ok);
// the same verdicts for k servers' proofs of the SAME n ciphertexts, records item-major (proof i*k + j), shared squarings
int zkp_verify_shared_dev(pgpu_ctx* ctx, size_t n, int k, const int* ids, const uint32_t* c, const uint32_t* dec, const uint32_t* e,
                          const uint32_t* z, uint8_t* ok);
int combine_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const uint32_t* decs, uint32_t* m_out, size_t share_stride = 0);
int combine_verified_dev(pgpu_ctx* ctx, size_t count, int k, const int* ids, const uint32_t* decs, size_t share_stride, const uint8_t* ok,
                         uint32_t* m_out, uint8_t* item_ok, size_t* n_failed, bool ok_item_major = false);   // ok[i*k + j] instead of ok[j*count + i]
int set_device(pgpu_ctx* ctx);

// ---- protocols.cu: level 2, alternative encryption, randomness extraction, nested operations, DDLEQ
void protocols_free(pgpu_ctx* ctx);
int fixed_table_build(pgpu_ctx* ctx, const ModCtx& M, const BigU& base, uint32_t exp_bits, FixedTable& T);
void fixed_table_free(FixedTable& T);
int modexp_fixed_dev(pgpu_ctx* ctx, const ModCtx& M, const FixedTable& T, size_t count, const ExpDesc& exp, uint32_t* out);
int ensure_fix_v(pgpu_ctx* ctx);
int setup_level2(pgpu_ctx* ctx);
int setup_level2_secret(pgpu_ctx* ctx);
int setup_level2_crt(pgpu_ctx* ctx);
int setup_alt(pgpu_ctx* ctx);
int encrypt2_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c);
int encrypt2_crt_dev(pgpu_ctx* ctx, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c);
int decrypt2_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, uint32_t* m);
int alt_encrypt_dev(pgpu_ctx* ctx, int level, size_t count, const uint32_t* m, const uint32_t* r, uint32_t* c);
int randomize_dev(pgpu_ctx* ctx, size_t count, const uint32_t* c, const uint32_t* r, uint32_t* out);
int extract_randomness_dev(pgpu_ctx* ctx, int level, size_t count, const uint32_t* c, uint32_t* out);
int nested_randomize_dev(pgpu_ctx* ctx, size_t count, const uint32_t* ct, const uint32_t* a, const uint32_t* b, uint32_t* out);
int ddleq_prove_dev(pgpu_ctx* ctx, size_t count, uint32_t secpar, const uint32_t* ct1, const uint32_t* ct2, const uint32_t* a, const uint32_t* b,
                    const uint32_t* x, const uint32_t* y, uint32_t* alpha, uint32_t* e, uint32_t* f, uint32_t* d_bad);
int ddleq_verify_dev(pgpu_ctx* ctx, size_t count, uint32_t secpar, const uint32_t* ct1, const uint32_t* ct2, const uint32_t* x, const uint32_t* y,
                     const uint32_t* alpha, const uint32_t* e, const uint32_t* f, uint8_t* ok);

}  // namespace pgpu
