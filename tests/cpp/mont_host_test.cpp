// Runs the multiplier SOURCES of the kernels -- paillier_b200/csrc/mont.cuh (32-bit limbs, integer pipe, incl. the dedicated
// squaring with its shared-memory exchange) and mont52.cuh (52-bit limbs as doubles, FP64 pipe) -- on the CPU, one thread per
// lane of an emulated warp (cuda_host_shim.h), and prints the results for tests/test_mont_host_emulation.py to compare with
// Python integers.  Built with -fsanitize=thread it is the race check of the squaring's exchange protocol (tools/host_sanitize.sh).
// stdin, whitespace separated, integers in hex:
//   m32 TPI L NSM  n  G  (a b) x G      -> G lines "mul sqr add sub"      (R = 2^(32*TPI*L); a, b < n; one (a, b) per lane group)
//   m52 TPI L S32  n  G  (a b) x G      -> G lines "mul add sub"          (R = 2^(52*TPI*L); records of S32 limbs)
//   m52s TPI L S32 lim  n  G  (a b) x G -> G lines "mul"                  (a read from, the product stored to, records of lim limbs)
#include "cuda_host_shim.h"

#include <iostream>
#include <string>

#include "../../paillier_b200/csrc/mont.cuh"
#include "../../paillier_b200/csrc/mont52.cuh"
#include "../../paillier_b200/csrc/vm_run.cuh"

using Limbs = std::vector<uint32_t>;

static Limbs parse_hex(std::string h, size_t limbs) {
    if (h.rfind("0x", 0) == 0) h = h.substr(2);
    Limbs v(limbs, 0);
    size_t nib = 0;
    for (size_t i = h.size(); i-- > 0; ++nib) {
        const char c = h[i];
        const uint32_t d = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
        if (nib / 8 >= limbs) { if (d) { std::cerr << "value wider than " << limbs << " limbs\n"; std::exit(2); } continue; }
        v[nib / 8] |= d << (4 * (nib % 8));
    }
    return v;
}
static std::string to_hex(const Limbs& v) {
    static const char* d = "0123456789abcdef";
    std::string s;
    for (size_t i = v.size(); i-- > 0;)
        for (int k = 7; k >= 0; --k) s += d[(v[i] >> (4 * k)) & 15];
    const size_t nz = s.find_first_not_of('0');
    return "0x" + (nz == std::string::npos ? std::string("0") : s.substr(nz));
}
static uint32_t neg_inv32(uint32_t n0) {
    uint32_t inv = n0;                                   // Newton: 3 correct bits, doubled four times
    for (int i = 0; i < 5; ++i) inv *= 2u - n0 * inv;
    return 0u - inv;
}

template <int TPI, int L, bool NSM>
static void run32(const Limbs& n, const std::vector<Limbs>& A, const std::vector<Limbs>& B) {
    using M = pgpu::Mont<TPI, L, NSM>;
    constexpr int S = TPI * L, G = 32 / TPI;
    std::vector<uint4> smem((size_t)M::SQR_ROWS * 32 + 1);
    std::vector<Limbs> mul(G, Limbs(S)), sqr(G, Limbs(S)), add(G, Limbs(S)), sub(G, Limbs(S));
    const uint32_t np0 = neg_inv32(n[0]);
    hostwarp::run_warp([&](int lane) {
        M m;
        m.init(n.data(), np0);
        if constexpr (M::HAS_SQR) m.init_sqr(smem.data());
        const int g = lane / TPI, t = lane % TPI;
        uint32_t a[L], b[L], r[L];
        for (int k = 0; k < L; ++k) { a[k] = A[g % A.size()][t * L + k]; b[k] = B[g % B.size()][t * L + k]; }
        m.mul(r, a, b);
        for (int k = 0; k < L; ++k) mul[g][t * L + k] = r[k];
        m.sqr(r, a);
        for (int k = 0; k < L; ++k) sqr[g][t * L + k] = r[k];
        m.add(r, a, b);
        for (int k = 0; k < L; ++k) add[g][t * L + k] = r[k];
        m.sub(r, a, b);
        for (int k = 0; k < L; ++k) sub[g][t * L + k] = r[k];
    });
    for (size_t g = 0; g < A.size() && g < (size_t)G; ++g)
        std::cout << to_hex(mul[g]) << " " << to_hex(sqr[g]) << " " << to_hex(add[g]) << " " << to_hex(sub[g]) << "\n";
}

template <int TPI, int L, int S32>
static void run52(const Limbs& n, const std::vector<Limbs>& A, const std::vector<Limbs>& B) {
    using M = pgpu::Mont52<TPI, L, S32>;
    constexpr int G = 32 / TPI;
    std::vector<Limbs> mul(G, Limbs(S32)), add(G, Limbs(S32)), sub(G, Limbs(S32));
    const uint32_t np0 = neg_inv32(n[0]);
    hostwarp::run_warp([&](int lane) {
        M m;
        m.init(n.data(), np0);
        const int g = lane / TPI;
        double x[L], y[L], r[L];
        m.load_rec(x, A[g % A.size()].data(), S32);
        m.load_rec(y, B[g % B.size()].data(), S32);
        m.mul(r, x, y);
        m.store_rec(mul[g].data(), r, S32);
        m.add(r, x, y);
        m.store_rec(add[g].data(), r, S32);
        m.sub(r, x, y);
        m.store_rec(sub[g].data(), r, S32);
    });
    for (size_t g = 0; g < A.size() && g < (size_t)G; ++g)
        std::cout << to_hex(mul[g]) << " " << to_hex(add[g]) << " " << to_hex(sub[g]) << "\n";
}

// short records: a is loaded from a record of `lim` limbs (EncryptWithR reads n-width m and r into the n^2 shape), the product is
// stored as a record of `lim` limbs (what an n-width output gets); the caller picks b so that the product fits
template <int TPI, int L, int S32>
static void run52_short(const Limbs& n, uint32_t lim, const std::vector<Limbs>& A, const std::vector<Limbs>& B) {
    using M = pgpu::Mont52<TPI, L, S32>;
    constexpr int G = 32 / TPI;
    std::vector<Limbs> mul(G, Limbs(S32, 0xdeadbeefu));
    const uint32_t np0 = neg_inv32(n[0]);
    hostwarp::run_warp([&](int lane) {
        M m;
        m.init(n.data(), np0);
        const int g = lane / TPI;
        double x[L], y[L], r[L];
        m.load_rec(x, A[g % A.size()].data(), lim);
        m.load_rec(y, B[g % B.size()].data(), S32);
        m.mul(r, x, y);
        m.store_rec(mul[g].data(), r, lim);
    });
    for (size_t g = 0; g < A.size() && g < (size_t)G; ++g) {
        for (uint32_t k = lim; k < (uint32_t)S32; ++k)
            if (mul[g][k] != 0xdeadbeefu) { std::cerr << "store_rec wrote past its " << lim << " limbs\n"; std::exit(3); }
        std::cout << to_hex(Limbs(mul[g].begin(), mul[g].begin() + lim)) << "\n";
    }
}

// ---- the interpreter itself (vm_run.cuh: what powm_vm / powm_vm52 wrap) on one emulated block of one warp: 32 / TPI resident
// groups, group g takes items g, g + n_groups, ...; a program the LIBRARY compiled (pgpu_selftest_last_program) is run for every
// item with the table, dump and constant buffers laid out as engine.cu's run_vm lays them out
struct VmJob {
    Limbs n, kconst, ops, in0, exps, out0, out1;
    uint32_t n_items = 0, in_limbs = 0, exp_stride = 0, exp_bits = 0, exp_sub = 0, out0_per_item = 1, tbl_entries = 1, flags = 0;
};
template <class B>
static void run_vm_job(VmJob& J) {
    constexpr int S = B::S32, TPI = B::TPI_, G = 32 / TPI, TS = TPI * B::TBL;
    J.out0.assign((size_t)J.n_items * J.out0_per_item * S, 0);
    J.out1.assign((size_t)J.n_items * S, 0);
    Limbs table((size_t)std::max<uint32_t>(J.tbl_entries, 1) * G * TS + 4, 0), dump((size_t)G * S, 0);
    std::vector<uint4> smem((size_t)pgpu::Mont<4, 8, true>::SQR_ROWS * 32 * 8 + 1);     // more than any shape's SQR_ROWS * 32
    hostwarp::shared_mem = smem.data();
    pgpu::VmParams P{};
    P.prog = J.ops.data(); P.n_items = J.n_items; P.mod = J.n.data(); P.np0 = neg_inv32(J.n[0]); P.kconst = J.kconst.data();
    for (int i = 0; i < pgpu::VM_MAX_IN; ++i) P.in_div[i] = 1;
    P.in[0] = J.in0.data(); P.in_stride[0] = J.in_limbs; P.in_limbs[0] = J.in_limbs;
    P.out[0] = J.out0.data(); P.out_stride[0] = J.out0_per_item * S; P.out_limbs[0] = S;
    P.out[1] = J.out1.data(); P.out_stride[1] = S; P.out_limbs[1] = S;
    P.exp = J.exps.data(); P.exp_stride = J.exp_stride; P.exp_bits = J.exp_bits; P.exp_sub = J.exp_sub;
    P.table = table.data(); P.n_groups = G; P.flags = J.flags; P.dump = dump.data();
    hostwarp::run_warp([&](int) { pgpu::vm_run<B>(P); });
}

// the shapes powm.cu builds (PGPU_FOR_EACH_SHAPE with NSM = SqrShape<TPI, L>::value, PGPU_FOR_EACH_SHAPE52)
#define SHAPES32(X) X(2, 16, false) X(4, 8, true) X(4, 16, true) X(8, 8, false) X(8, 12, false) X(4, 24, false) X(8, 16, false) \
                    X(16, 8, false) X(4, 32, false) X(32, 4, false) X(8, 24, false) X(16, 12, false) X(32, 6, false)
#define SHAPES52(X) X(4, 5, 32) X(4, 10, 64) X(8, 5, 64) X(4, 15, 96) X(8, 8, 96) X(8, 10, 128) X(8, 15, 192) X(16, 8, 192)
// the interpreter is instantiated for a subset (compile time): both squaring shapes, the widest integer shape, two FP64 shapes
// incl. the one that serves 4096-bit moduli
#define VM_SHAPES32(X) X(4, 8, true) X(4, 16, true) X(8, 8, false) X(4, 32, false)
#define VM_SHAPES52(X) X(4, 5, 32) X(8, 10, 128)

int main() {
    std::string kind;
    while (std::cin >> kind) {
        if (kind == "vm32" || kind == "vm52") {
            // vm32 TPI L | vm52 TPI L S32 ; flags n_items in_limbs exp_stride exp_bits exp_sub out0_per_item tbl_entries ; n ; R1 R2 ;
            // n_ops ops... ; per item: base exps(one number of exp_stride limbs)   ->  per item: out0 records..., out1
            int tpi, l, s32 = 0;
            std::cin >> tpi >> l;
            if (kind == "vm52") std::cin >> s32; else s32 = tpi * l;
            VmJob J;
            size_t n_ops;
            std::string nh, r1, r2;
            std::cin >> J.flags >> J.n_items >> J.in_limbs >> J.exp_stride >> J.exp_bits >> J.exp_sub >> J.out0_per_item >> J.tbl_entries >> nh >> r1 >> r2 >> n_ops;
            J.n = parse_hex(nh, s32);
            J.kconst.assign((size_t)16 * s32, 0);                                   // engine.hpp: K_R2 = 0, K_R1 = 1, K_ONE = 2 of K_SLOTS = 16
            const Limbs R2 = parse_hex(r2, s32), R1 = parse_hex(r1, s32);
            std::copy(R2.begin(), R2.end(), J.kconst.begin());
            std::copy(R1.begin(), R1.end(), J.kconst.begin() + s32);
            J.kconst[2 * (size_t)s32] = 1;
            for (size_t i = 0; i < n_ops; ++i) { std::string o; std::cin >> o; J.ops.push_back((uint32_t)std::stoul(o, nullptr, 16)); }
            for (uint32_t i = 0; i < J.n_items; ++i) {
                std::string b, e; std::cin >> b >> e;
                const Limbs bl = parse_hex(b, J.in_limbs), el = parse_hex(e, std::max<uint32_t>(J.exp_stride, 1));
                J.in0.insert(J.in0.end(), bl.begin(), bl.end());
                J.exps.insert(J.exps.end(), el.begin(), el.begin() + J.exp_stride);
            }
            J.in0.resize(J.in0.size() + 8, 0); J.exps.resize(J.exps.size() + 8, 0);
            bool done = false;
            if (kind == "vm32") {
#define X(T, LL, NSM) if (!done && tpi == T && l == LL) { run_vm_job<pgpu::Vm32<T, LL>>(J); done = true; }
                VM_SHAPES32(X)
#undef X
            } else {
#define X(T, LL, SS) if (!done && tpi == T && l == LL && s32 == SS) { run_vm_job<pgpu::Vm52<T, LL, SS>>(J); done = true; }
                VM_SHAPES52(X)
#undef X
            }
            if (!done) { std::cerr << "shape not built: " << kind << " " << tpi << " " << l << "\n"; return 2; }
            for (uint32_t i = 0; i < J.n_items; ++i) {
                for (uint32_t j = 0; j < J.out0_per_item; ++j) {
                    const size_t o = ((size_t)i * J.out0_per_item + j) * s32;
                    std::cout << to_hex(Limbs(J.out0.begin() + o, J.out0.begin() + o + s32)) << " ";
                }
                std::cout << to_hex(Limbs(J.out1.begin() + (size_t)i * s32, J.out1.begin() + (size_t)(i + 1) * s32)) << "\n";
            }
            continue;
        }
        int tpi, l, third;
        uint32_t lim = 0;
        std::string nh;
        size_t groups;
        std::cin >> tpi >> l >> third;
        if (kind == "m52s") std::cin >> lim;
        std::cin >> nh >> groups;
        const size_t limbs = kind == "m32" ? (size_t)tpi * l : (size_t)third;
        const Limbs n = parse_hex(nh, limbs);
        std::vector<Limbs> A, B;
        for (size_t g = 0; g < groups; ++g) {
            std::string a, b; std::cin >> a >> b;
            A.push_back(parse_hex(a, limbs));
            B.push_back(parse_hex(b, limbs));
            if (kind == "m52s")                      // what lies behind a short record (the next item's record) must not be read
                for (size_t k = lim; k < limbs; ++k) A.back()[k] = 0xa5a5a5a5u;
        }
        bool done = false;
        if (kind == "m32") {
#define X(T, LL, NSM) if (!done && tpi == T && l == LL && (third != 0) == NSM) { static_assert(pgpu::SqrShape<T, LL>::value == NSM, "NSM follows SqrShape"); run32<T, LL, NSM>(n, A, B); done = true; }
            SHAPES32(X)
#undef X
        } else if (kind == "m52") {
#define X(T, LL, SS) if (!done && tpi == T && l == LL && third == SS) { run52<T, LL, SS>(n, A, B); done = true; }
            SHAPES52(X)
#undef X
        } else if (kind == "m52s") {
#define X(T, LL, SS) if (!done && tpi == T && l == LL && third == SS) { run52_short<T, LL, SS>(n, lim, A, B); done = true; }
            SHAPES52(X)
#undef X
        }
        if (!done) { std::cerr << "shape not built: " << kind << " " << tpi << " " << l << " " << third << "\n"; return 2; }
    }
    return 0;
}
