// Launch interface of aux_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pgpu {

constexpr int CRT_MAXH = 128;  // limbs of p, q (CRT recombination) and of n (share combining) supported

// consts: 7 records of h limbs: p, q, p^-1 mod 2^(32h), q^-1 mod 2^(32h),
// h_p*2^(32h) mod p, h_q*2^(32h) mod q, q^-1*2^(32h) mod p
struct CrtParams {
    uint32_t n_items;
    int h;
    const uint32_t* consts;
    uint32_t np0_p, np0_q;
    const uint32_t* xp;     // c^(p-1) mod p^2, x_stride limbs apart
    const uint32_t* xq;     // c^(q-1) mod q^2
    uint32_t x_stride;
    uint32_t* out;          // plaintexts, out_stride limbs apart, out_limbs written
    uint32_t out_stride;
    uint32_t out_limbs;
};

cudaError_t crt_combine_launch(const CrtParams& P, cudaStream_t stream);

// out (S limbs) = product of n_items records (stride S) times 2^(-32*S*T) mod N,
// T = number of Montgomery multiplications performed (returned by prod_reduce_mults()).
struct ProdParams {
    const uint32_t* in;      // n_items records of S limbs
    uint32_t n_items;
    const uint32_t* mod;
    uint32_t np0;
    uint32_t* partial;       // one record per block
};
cudaError_t prod_reduce_launch(int tpi, int limbs, const ProdParams& P, int blocks, cudaStream_t stream);

// m = L(c') * K mod n with K = (4*delta^2)^-1 mod n (thresholdkey.go:143-146, :63-66).
// consts: 3 records of h limbs: n, n^-1 mod 2^(32h), K*2^(32h) mod n
struct CombineParams {
    uint32_t n_items;
    int h;
    const uint32_t* consts;
    uint32_t np0;
    const uint32_t* cprime;  // records of cp_stride limbs (n^2 width)
    uint32_t cp_stride;
    uint32_t* out;           // h limbs per item
};
cudaError_t combine_final_launch(const CombineParams& P, cudaStream_t stream);

// Level-2 recovery of Decrypt (recoveryAlgorithm for s = 2, paillier.go:308-340) followed by the
// multiplication with lambda^-1 mod n^2 (:298-300).  tmp = c^lambda mod n^3 in records of t_stride limbs.
// consts (H = 2h limbs each): n^2, n^-1 mod 2^(32H), R^2 mod n^2 (R = 2^(32H)), inv2*R mod n^2,
// mu*R mod n^2 (mu = lambda^-1 mod n^2), n*R mod n^2
struct Recover2Params {
    uint32_t n_items;
    int h;                   // limbs of n; n^2 arithmetic uses H = 2h limbs
    const uint32_t* consts;
    uint32_t np0_n2;
    const uint32_t* tmp; uint32_t t_stride; uint32_t t_limbs;
    uint32_t* out; uint32_t out_stride;   // H limbs written
};
cudaError_t recover2_launch(const Recover2Params& P, cudaStream_t stream);

// Decrypt at level 2 by CRT over p^3 and q^3.  xp = c^(p-1) mod p^3, xq = c^(q-1) mod q^3 (records of x_stride limbs).
// With e = m(p-1) mod p^2:  (xp - 1)/p = e q + C(e,2) p q^2 (mod p^2); e mod p comes from the residue mod p, the binomial
// term is then known, and m mod p^2 = (L1 - p (C(e0,2) q^2 mod p)) (q (p-1))^-1 mod p^2.  Garner over p^2, q^2 gives m mod n^2
// (= recoveryAlgorithm(c^lambda mod n^3, 2) * lambda^-1 mod n^2, paillier.go:292-340).
// consts per prime (h = limbs of the prime, H = 2h): P2 (H), p^-1 mod 2^(32H) (H), Cm*R mod p^2 (H), p (h), R_h^2 mod p (h),
// R_h^3 mod p (h), q^-1 R_h mod p (h), inv2 q^2 R_h mod p (h), R_h mod p (h); then for Garner: (q^2)^-1 R mod p^2 (H), q^2 (H).
struct Crt2Params {
    uint32_t n_items;
    int h;
    const uint32_t* cp;      // constants of p (other prime q)
    const uint32_t* cq;      // constants of q (other prime p)
    const uint32_t* garner;  // (q^2)^-1 * R mod p^2, q^2
    uint32_t np0_p, np0_p2, np0_q, np0_q2;
    const uint32_t* xp; const uint32_t* xq; uint32_t x_stride;
    uint32_t* out;           // out_limbs (<= 2H) limbs per item
    uint32_t out_limbs;
    uint32_t* mp; uint32_t* mq;   // scratch: m mod p^2, m mod q^2 (2h limbs per item each)
};
cudaError_t crt2_launch(const Crt2Params& P, cudaStream_t stream);

// ---- bigops.cu ----
constexpr int BIG_MAXS = 192;

struct InvParams {
    uint32_t n_items;
    int limbs;
    const uint32_t* mod;
    const uint32_t* in;
    uint32_t* out;
    uint32_t* first_bad;     // atomicMin of the first non-invertible item index
};
cudaError_t modinv_launch(const InvParams& P, cudaStream_t stream);

struct MulParams {
    uint32_t n_items;
    const uint32_t* a; uint32_t a_stride; int na;
    const uint32_t* b; uint32_t b_stride; int nb;
    uint32_t* out; uint32_t out_stride; uint32_t out_limbs;
};
cudaError_t bigmul_launch(const MulParams& P, cudaStream_t stream);

struct MulAddParams {
    uint32_t n_items;
    const uint32_t* r; uint32_t r_stride; uint32_t nr;
    const uint32_t* e; uint32_t e_stride; int ne;
    const uint32_t* k; int nk;
    uint32_t* out; uint32_t out_stride; uint32_t out_limbs;
};
cudaError_t muladd_launch(const MulAddParams& P, cudaStream_t stream);

struct ShaParams {
    uint32_t n_items;
    int n_seg;
    const uint32_t* seg[6];
    uint32_t stride[6];      // limbs between items (0 = same value for every item)
    int limbs[6];
    uint32_t div[6];         // item i hashes record i / div (0 is read as 1)
    uint32_t* out;           // 8 limbs per item
};
cudaError_t sha256_concat_launch(const ShaParams& P, cudaStream_t stream);

cudaError_t equal_launch(const uint32_t* a, const uint32_t* b, uint32_t limbs, uint32_t n_items, uint8_t* flags, cudaStream_t stream);
// *first = smallest i with flags[i] == 0, else 0xffffffff
cudaError_t first_zero_launch(const uint8_t* flags, uint32_t n_items, uint32_t* first, cudaStream_t stream);
// out[i] = (digest[i] & 1) ? a[i / a_div] : b[i / b_div]
cudaError_t select_launch(const uint32_t* digest, const uint32_t* a, uint32_t a_div, const uint32_t* b, uint32_t b_div, uint32_t limbs, uint32_t n_items, uint32_t* out, cudaStream_t stream);
// out[i] = in[i / rep] (records of `limbs` limbs): a statement value repeated for each of its proof instances
cudaError_t repeat_launch(const uint32_t* in, uint32_t limbs, uint32_t rep, uint32_t* out, uint32_t n_out, cudaStream_t stream);
// out[i] = in[idx[i] / div] and out[idx[i]] = in[i] over records of `limbs` limbs (compaction of the DDLEQ instances whose
// challenge bit is 1)
cudaError_t gather_launch(const uint32_t* in, uint32_t limbs, const uint32_t* idx, uint32_t div, uint32_t* out, uint32_t n, cudaStream_t stream);
cudaError_t scatter_launch(const uint32_t* in, uint32_t limbs, const uint32_t* idx, uint32_t* out, uint32_t n, cudaStream_t stream);
cudaError_t resize_launch(const uint32_t* in, uint32_t in_stride, uint32_t in_limbs, uint32_t* out, uint32_t out_limbs, uint32_t n_items, cudaStream_t stream);

}  // namespace pgpu
