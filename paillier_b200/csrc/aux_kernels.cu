// Light per-item epilogues and reductions around powm_vm.  They are <1 % of a
// batch's multiply work, so they use plain one-thread-per-item limb loops
// (the CRT recombination) or the warp-cooperative multiplier on a tree (the
// Add reduction).
#include <cuda_runtime.h>
#include <cstdint>
#include "mont.cuh"
#include "aux.h"

namespace pgpu {

// ---------------------------------------------------------------------------
// one-thread helpers on little-endian limb arrays of h limbs (h <= CRT_MAXH)
// ---------------------------------------------------------------------------
// r = a * b * 2^(-32h) mod n   (CIOS, n odd, a < 2^(32h), b < n)
__device__ void st_mont(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* n, uint32_t np0, int h) {
    uint32_t t[CRT_MAXH + 2];
    for (int i = 0; i < h + 2; ++i) t[i] = 0;
    for (int i = 0; i < h; ++i) {
        uint64_t c = 0;
        const uint32_t bi = b[i];
        for (int j = 0; j < h; ++j) {
            c += (uint64_t)a[j] * bi + t[j];
            t[j] = (uint32_t)c; c >>= 32;
        }
        c += t[h]; t[h] = (uint32_t)c; t[h + 1] = (uint32_t)(c >> 32);
        const uint32_t q = t[0] * np0;
        c = ((uint64_t)n[0] * q + t[0]) >> 32;
        for (int j = 1; j < h; ++j) {
            c += (uint64_t)n[j] * q + t[j];
            t[j - 1] = (uint32_t)c; c >>= 32;
        }
        c += t[h]; t[h - 1] = (uint32_t)c;
        t[h] = t[h + 1] + (uint32_t)(c >> 32);
    }
    // conditional subtraction
    bool ge = t[h] != 0;
    if (!ge) {
        ge = true;
        for (int j = h - 1; j >= 0; --j) {
            if (t[j] != n[j]) { ge = t[j] > n[j]; break; }
        }
    }
    if (ge) {
        int64_t bw = 0;
        for (int j = 0; j < h; ++j) {
            int64_t d = (int64_t)t[j] - n[j] - bw;
            bw = d < 0; t[j] = (uint32_t)d;
        }
    }
    for (int j = 0; j < h; ++j) r[j] = t[j];
}

// Paillier's L over one prime: k = (x - 1) / p for x = 1 (mod p), x < p^2,
// as an exact division: k = (x - 1) * p^-1 mod 2^(32h)   (paillier.go:437-440)
__device__ void st_L(uint32_t* k, const uint32_t* x, const uint32_t* pinv, int h) {
    uint32_t y[CRT_MAXH];
    int64_t bw = 1;
    for (int j = 0; j < h; ++j) {
        int64_t d = (int64_t)x[j] - bw;
        bw = d < 0; y[j] = (uint32_t)d;
    }
    for (int j = 0; j < h; ++j) k[j] = 0;
    for (int i = 0; i < h; ++i) {
        uint64_t c = 0;
        const uint32_t yi = y[i];
        for (int j = 0; i + j < h; ++j) {
            c += (uint64_t)yi * pinv[j] + k[i + j];
            k[i + j] = (uint32_t)c; c >>= 32;
        }
    }
}

// m = CRT(m_p, m_q) with m_p = L_p(x_p) * h_p mod p, m_q = L_q(x_q) * h_q mod q
// where x_p = c^(p-1) mod p^2, x_q = c^(q-1) mod q^2.  Equal to the reference's
// L(c^lambda mod n^2) * lambda^-1 mod n (paillier.go:292-303) for every valid ciphertext.
__global__ void crt_combine_kernel(CrtParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_items) return;
    const int h = P.h;
    const uint32_t* K = P.consts;
    const uint32_t* p = K + 0 * h;  const uint32_t* q = K + 1 * h;
    const uint32_t* pinv = K + 2 * h; const uint32_t* qinv = K + 3 * h;
    const uint32_t* hpM = K + 4 * h; const uint32_t* hqM = K + 5 * h;
    const uint32_t* cqM = K + 6 * h;
    uint32_t mp[CRT_MAXH], mq[CRT_MAXH], a[CRT_MAXH], b[CRT_MAXH];
    st_L(a, P.xp + (size_t)i * P.x_stride, pinv, h);
    st_mont(mp, a, hpM, p, P.np0_p, h);
    st_L(a, P.xq + (size_t)i * P.x_stride, qinv, h);
    st_mont(mq, a, hqM, q, P.np0_q, h);
    // t = (m_p - m_q) * q^-1 mod p
    st_mont(a, mp, cqM, p, P.np0_p, h);
    st_mont(b, mq, cqM, p, P.np0_p, h);
    int64_t bw = 0;
    for (int j = 0; j < h; ++j) {
        int64_t d = (int64_t)a[j] - b[j] - bw;
        bw = d < 0; a[j] = (uint32_t)d;
    }
    if (bw) {
        uint64_t c = 0;
        for (int j = 0; j < h; ++j) { c += (uint64_t)a[j] + p[j]; a[j] = (uint32_t)c; c >>= 32; }
    }
    // m = m_q + q * t
    uint32_t r[2 * CRT_MAXH];
    for (int j = 0; j < 2 * h; ++j) r[j] = j < h ? mq[j] : 0;
    for (int ii = 0; ii < h; ++ii) {
        uint64_t c = 0;
        const uint32_t ti = a[ii];
        for (int j = 0; j < h; ++j) {
            c += (uint64_t)ti * q[j] + r[ii + j];
            r[ii + j] = (uint32_t)c; c >>= 32;
        }
        for (int j = ii + h; c != 0 && j < 2 * h; ++j) {
            c += r[j]; r[j] = (uint32_t)c; c >>= 32;
        }
    }
    uint32_t* out = P.out + (size_t)i * P.out_stride;
    for (uint32_t j = 0; j < P.out_limbs; ++j) out[j] = j < (uint32_t)(2 * h) ? r[j] : 0;
}

cudaError_t crt_combine_launch(const CrtParams& P, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    const int threads = 64;
    crt_combine_kernel<<<(P.n_items + threads - 1) / threads, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

// m = L(c') * (4*delta^2)^-1 mod n   (computeDecryption, thresholdkey.go:143-146)
__global__ void combine_final_kernel(CombineParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_items) return;
    const int h = P.h;
    const uint32_t* n = P.consts;
    const uint32_t* ninv = P.consts + h;
    const uint32_t* KM = P.consts + 2 * h;
    uint32_t l[CRT_MAXH], m[CRT_MAXH];
    st_L(l, P.cprime + (size_t)i * P.cp_stride, ninv, h);
    st_mont(m, l, KM, n, P.np0, h);
    uint32_t* out = P.out + (size_t)i * h;
    for (int j = 0; j < h; ++j) out[j] = m[j];
}

cudaError_t combine_final_launch(const CombineParams& P, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    const int threads = 64;
    combine_final_kernel<<<(P.n_items + threads - 1) / threads, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

// r = (a + b) mod n or (a - b) mod n on H-limb arrays, a, b < n
__device__ void st_addmod(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* n, int H) {
    uint64_t c = 0;
    for (int j = 0; j < H; ++j) { c += (uint64_t)a[j] + b[j]; r[j] = (uint32_t)c; c >>= 32; }
    bool ge = c != 0;
    if (!ge) { ge = true; for (int j = H - 1; j >= 0; --j) if (r[j] != n[j]) { ge = r[j] > n[j]; break; } }
    if (ge) { int64_t bw = 0; for (int j = 0; j < H; ++j) { int64_t d = (int64_t)r[j] - n[j] - bw; bw = d < 0; r[j] = (uint32_t)d; } }
}
__device__ void st_submod(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* n, int H) {
    int64_t bw = 0;
    for (int j = 0; j < H; ++j) { int64_t d = (int64_t)a[j] - b[j] - bw; bw = d < 0; r[j] = (uint32_t)d; }
    if (bw) { uint64_t c = 0; for (int j = 0; j < H; ++j) { c += (uint64_t)r[j] + n[j]; r[j] = (uint32_t)c; c >>= 32; } }
}

// Decrypt at level 2: m = recoveryAlgorithm(c^lambda mod n^3, 2) * lambda^-1 mod n^2  (paillier.go:298-340)
__global__ void recover2_kernel(Recover2Params P) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P.n_items) return;
    const int h = P.h, H = 2 * h;
    const uint32_t* N2 = P.consts;            const uint32_t* ninv = P.consts + H;
    const uint32_t* R2 = P.consts + 2 * H;    const uint32_t* inv2M = P.consts + 3 * H;
    const uint32_t* muM = P.consts + 4 * H;   const uint32_t* nM = P.consts + 5 * H;
    const uint32_t* tmp = P.tmp + (size_t)idx * P.t_stride;
    uint32_t lo[CRT_MAXH], hi[CRT_MAXH], one[CRT_MAXH], a[CRT_MAXH], b[CRT_MAXH], t1[CRT_MAXH];
    for (int j = 0; j < H; ++j) {
        lo[j] = tmp[j];
        hi[j] = (uint32_t)(H + j) < P.t_limbs ? tmp[H + j] : 0;
        one[j] = j == 0;
    }
    // j = 1: amod = tmp mod n^2 (:316); i1 = L(amod, n) (:318), an h-limb value
    st_mont(a, lo, R2, N2, P.np0_n2, H);      // lo * R
    st_mont(a, a, one, N2, P.np0_n2, H);      // lo mod n^2
    st_mont(b, hi, R2, N2, P.np0_n2, H);      // hi * R mod n^2
    st_addmod(a, a, b, N2, H);                // tmp mod n^2
    uint32_t i1[CRT_MAXH];
    st_L(i1, a, ninv, h);                     // (amod - 1) / n, exact
    for (int j = h; j < H; ++j) i1[j] = 0;
    // j = 2: t1 = L(tmp mod n^3, n) = (tmp - 1) / n (:316-318), an H-limb value
    st_L(t1, tmp, ninv, H);
    // k = 2 (:321-334): t2 = i1*(i1-1) mod n^2; t2 = t2 * n * 2!^-1; t1 = (t1 - t2) mod n^2
    bool i1_zero = true;
    for (int j = 0; j < h; ++j) i1_zero = i1_zero && i1[j] == 0;
    if (!i1_zero) {
        int64_t bw = 1;
        for (int j = 0; j < H; ++j) { int64_t d = (int64_t)i1[j] - bw; bw = d < 0; b[j] = (uint32_t)d; }   // i1 - 1
        st_mont(a, i1, R2, N2, P.np0_n2, H);          // i1 * R
        st_mont(a, a, b, N2, P.np0_n2, H);            // i1 * (i1 - 1) mod n^2
        st_mont(a, a, nM, N2, P.np0_n2, H);           // * n
        st_mont(a, a, inv2M, N2, P.np0_n2, H);        // * 2^-1
        st_submod(t1, t1, a, N2, H);
    }
    // m = i2 * mu mod n^2 (:298-300)
    st_mont(a, t1, muM, N2, P.np0_n2, H);
    uint32_t* out = P.out + (size_t)idx * P.out_stride;
    for (int j = 0; j < H; ++j) out[j] = a[j];
}

cudaError_t recover2_launch(const Recover2Params& P, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    const int threads = 32;
    recover2_kernel<<<(P.n_items + threads - 1) / threads, threads, 0, stream>>>(P);
    return cudaGetLastError();
}


// m mod P^2 for one prime P of the level-2 CRT decryption (see Crt2Params)
__device__ void crt2_half(uint32_t* mP, const uint32_t* x, const uint32_t* K, uint32_t np0_p, uint32_t np0_p2, int h) {
    const int H = 2 * h;
    const uint32_t* P2 = K;            const uint32_t* pinvH = K + H;     const uint32_t* CmM = K + 2 * H;
    const uint32_t* p = K + 3 * H;     const uint32_t* R2h = p + h;       const uint32_t* R3h = p + 2 * h;
    const uint32_t* CqM = p + 3 * h;   const uint32_t* C2M = p + 4 * h;   const uint32_t* oneM = p + 5 * h;
    uint32_t L1[CRT_MAXH], a[CRT_MAXH], b[CRT_MAXH], w[CRT_MAXH];
    st_L(L1, x, pinvH, H);                              // (x - 1) / p, exact, < p^2
    // e0 = (L1 mod p) * q^-1 mod p, Montgomery form (R_h)
    st_mont(a, L1, R2h, p, np0_p, h);                   // l0 * R_h
    st_mont(b, L1 + h, R3h, p, np0_p, h);               // l1 * R_h^2
    st_addmod(a, a, b, p, h);                           // (L1 mod p) * R_h
    st_mont(a, a, CqM, p, np0_p, h);                    // e0 * R_h
    // w = C(e0, 2) * q^2 mod p, plain
    st_submod(b, a, oneM, p, h);                        // (e0 - 1) * R_h
    st_mont(b, a, b, p, np0_p, h);                      // e0 (e0 - 1) * R_h
    st_mont(b, b, C2M, p, np0_p, h);                    // * inv2 * q^2
    for (int j = 0; j < h; ++j) a[j] = j == 0;
    st_mont(w, b, a, p, np0_p, h);                      // out of Montgomery form
    // V = L1 - p * w mod p^2
    for (int j = 0; j < H; ++j) a[j] = 0;
    for (int i = 0; i < h; ++i) {
        uint64_t c = 0;
        const uint32_t wi = w[i];
        for (int j = 0; j < h; ++j) { c += (uint64_t)wi * p[j] + a[i + j]; a[i + j] = (uint32_t)c; c >>= 32; }
        for (int j = i + h; c != 0 && j < H; ++j) { c += a[j]; a[j] = (uint32_t)c; c >>= 32; }
    }
    st_submod(b, L1, a, P2, H);
    st_mont(mP, b, CmM, P2, np0_p2, H);                 // * (q (p-1))^-1 mod p^2
}

// one prime per launch: m mod P^2 into mhalf (H limbs per item)
__global__ void crt2_half_kernel(uint32_t n_items, int h, const uint32_t* consts, uint32_t np0_p, uint32_t np0_p2,
                                 const uint32_t* x, uint32_t x_stride, uint32_t* mhalf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    uint32_t m[CRT_MAXH];
    crt2_half(m, x + (size_t)i * x_stride, consts, np0_p, np0_p2, h);
    uint32_t* out = mhalf + (size_t)i * 2 * h;
    for (int j = 0; j < 2 * h; ++j) out[j] = m[j];
}

// Garner over the moduli p^2, q^2: m = mq + q^2 * ((mp - mq) * (q^2)^-1 mod p^2)
__global__ void crt2_garner_kernel(Crt2Params P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_items) return;
    const int H = 2 * P.h;
    const uint32_t* mp = P.mp + (size_t)i * H;
    const uint32_t* mq = P.mq + (size_t)i * H;
    const uint32_t* P2 = P.cp;
    const uint32_t* CgM = P.garner;
    const uint32_t* Q2 = P.garner + H;
    uint32_t a[CRT_MAXH], b[CRT_MAXH];
    st_mont(a, mp, CgM, P2, P.np0_p2, H);
    st_mont(b, mq, CgM, P2, P.np0_p2, H);
    st_submod(a, a, b, P2, H);
    uint32_t* out = P.out + (size_t)i * P.out_limbs;
    uint32_t r[2 * CRT_MAXH];
    for (int j = 0; j < 2 * H; ++j) r[j] = j < H ? mq[j] : 0;
    for (int ii = 0; ii < H; ++ii) {
        uint64_t c = 0;
        const uint32_t ti = a[ii];
        for (int j = 0; j < H; ++j) { c += (uint64_t)ti * Q2[j] + r[ii + j]; r[ii + j] = (uint32_t)c; c >>= 32; }
        for (int j = ii + H; c != 0 && j < 2 * H; ++j) { c += r[j]; r[j] = (uint32_t)c; c >>= 32; }
    }
    for (uint32_t j = 0; j < P.out_limbs; ++j) out[j] = r[j];
}

cudaError_t crt2_launch(const Crt2Params& P, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    const int threads = 32;
    const unsigned blocks = (P.n_items + threads - 1) / threads;
    crt2_half_kernel<<<blocks, threads, 0, stream>>>(P.n_items, P.h, P.cp, P.np0_p, P.np0_p2, P.xp, P.x_stride, P.mp);
    crt2_half_kernel<<<blocks, threads, 0, stream>>>(P.n_items, P.h, P.cq, P.np0_q, P.np0_q2, P.xq, P.x_stride, P.mq);
    crt2_garner_kernel<<<blocks, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// PublicKey.Add over a batch (operations.go:11-29): one running Montgomery
// product per group, then a shared-memory tree across the block's groups.
// Missing items are padded with the plain integer 1, so every block performs
// exactly GPB*rounds - 1 multiplications and the power of R^-1 picked up on the
// way is known to the host, which folds the correction into a final multiply.
// ---------------------------------------------------------------------------
template <int TPI, int L>
__global__ void __launch_bounds__(128) prod_reduce_kernel(ProdParams P) {
    constexpr int S = TPI * L;
    constexpr int GPB = 128 / TPI;
    __shared__ uint32_t sm[GPB * S];
    Mont<TPI, L> M;
    M.init(P.mod, P.np0);
    const int g = threadIdx.x / TPI, t = threadIdx.x & (TPI - 1);
    const uint32_t G = gridDim.x * GPB, group = blockIdx.x * GPB + g;
    const uint32_t rounds = P.n_items == 0 ? 1 : (P.n_items + G - 1) / G;
    uint32_t acc[L], y[L];
    for (uint32_t rd = 0; rd < rounds; ++rd) {
        const uint32_t idx = rd * G + group;
        if (idx < P.n_items) {
            const uint32_t* p = P.in + (size_t)idx * S + t * L;
#pragma unroll
            for (int k = 0; k < L; ++k) y[k] = __ldg(p + k);
        } else {
#pragma unroll
            for (int k = 0; k < L; ++k) y[k] = (t == 0 && k == 0) ? 1u : 0u;
        }
        if (rd == 0) {
#pragma unroll
            for (int k = 0; k < L; ++k) acc[k] = y[k];
        } else {
            M.mul(acc, acc, y);
        }
    }
    for (int stride = GPB / 2; stride >= 1; stride >>= 1) {
#pragma unroll
        for (int k = 0; k < L; ++k) sm[g * S + t * L + k] = acc[k];
        __syncthreads();
        const int partner = (g + stride) % GPB;
#pragma unroll
        for (int k = 0; k < L; ++k) y[k] = sm[partner * S + t * L + k];
        __syncthreads();
        M.mul(acc, acc, y);
    }
    if (g == 0) {
        uint32_t* o = P.partial + (size_t)blockIdx.x * S + t * L;
#pragma unroll
        for (int k = 0; k < L; ++k) o[k] = acc[k];
    }
}

#define PGPU_PROD_SHAPES(X) X(4, 8) X(4, 16) X(8, 12) X(4, 24) X(8, 16) X(4, 32) X(8, 24) X(16, 12)

cudaError_t prod_reduce_launch(int tpi, int limbs, const ProdParams& P, int blocks, cudaStream_t stream) {
#define X(T, LL) if (tpi == T && limbs == LL) { prod_reduce_kernel<T, LL><<<blocks, 128, 0, stream>>>(P); return cudaGetLastError(); }
    PGPU_PROD_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace pgpu
