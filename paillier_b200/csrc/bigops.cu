// One-thread-per-item big-integer kernels for the light steps around the
// exponentiations: modular inverse (gmp ModInverse call sites),
// unreduced products (c^4 and c_i^2 of the ZKP transcript, thresholdkey.go:241,248),
// Z = r + E*delta*share (thresholdkey.go:313-317), SHA-256 over minimal
// big-endian Bytes() (thresholdkey.go:319-326, random_oracle.go:20-32) and the
// L(c')*const mod n tail of share combining (thresholdkey.go:143-146).
// Together they are well below 1 % of a batch's multiply work.
#include <cuda_runtime.h>
#include <cstdint>
#include "aux.h"

namespace pgpu {

// ------------------------------------------------------------------ modinv
__device__ __forceinline__ bool bn_is_one(const uint32_t* a, int n) {
    if (a[0] != 1) return false;
    for (int i = 1; i < n; ++i) if (a[i]) return false;
    return true;
}
__device__ __forceinline__ bool bn_is_zero(const uint32_t* a, int n) {
    for (int i = 0; i < n; ++i) if (a[i]) return false;
    return true;
}
__device__ __forceinline__ int bn_cmp(const uint32_t* a, const uint32_t* b, int n) {
    for (int i = n - 1; i >= 0; --i) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}
__device__ __forceinline__ void bn_shr1(uint32_t* a, int n, uint32_t top) {
    for (int i = 0; i < n - 1; ++i) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[n - 1] = (a[n - 1] >> 1) | (top << 31);
}
__device__ __forceinline__ uint32_t bn_add(uint32_t* a, const uint32_t* b, int n) {   // a += b, returns carry
    uint64_t c = 0;
    for (int i = 0; i < n; ++i) { c += (uint64_t)a[i] + b[i]; a[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
__device__ __forceinline__ uint32_t bn_sub(uint32_t* a, const uint32_t* b, int n) {   // a -= b, returns borrow
    int64_t bw = 0;
    for (int i = 0; i < n; ++i) { int64_t d = (int64_t)a[i] - b[i] - bw; bw = d < 0; a[i] = (uint32_t)d; }
    return (uint32_t)bw;
}
// x = x / 2 mod m (m odd)
__device__ __forceinline__ void bn_half_mod(uint32_t* x, const uint32_t* m, int n) {
    uint32_t top = 0;
    if (x[0] & 1) top = bn_add(x, m, n);
    bn_shr1(x, n, top);
}
// x = (x - y) mod m, x, y < m
__device__ __forceinline__ void bn_sub_mod(uint32_t* x, const uint32_t* y, const uint32_t* m, int n) {
    if (bn_sub(x, y, n)) bn_add(x, m, n);
}

// out = a^-1 mod m by the binary extended Euclid (m odd); flag = 1 if gcd(a, m) != 1
__global__ void modinv_kernel(InvParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_items) return;
    const int n = P.limbs;
    uint32_t u[BIG_MAXS], v[BIG_MAXS], x1[BIG_MAXS], x2[BIG_MAXS], m[BIG_MAXS];
    const uint32_t* a = P.in + (size_t)i * P.limbs;
    for (int k = 0; k < n; ++k) { m[k] = P.mod[k]; u[k] = a[k]; v[k] = m[k]; x1[k] = 0; x2[k] = 0; }
    x1[0] = 1;
    // reduce a mod m if needed (inputs are canonical in every caller; a >= m only by misuse)
    while (bn_cmp(u, m, n) >= 0) bn_sub(u, m, n);
    bool ok = !bn_is_zero(u, n);
    while (ok && !bn_is_one(u, n) && !bn_is_one(v, n)) {
        while (!(u[0] & 1)) { bn_shr1(u, n, 0); bn_half_mod(x1, m, n); }
        while (!(v[0] & 1)) { bn_shr1(v, n, 0); bn_half_mod(x2, m, n); }
        const int c = bn_cmp(u, v, n);
        if (c == 0) { ok = bn_is_one(u, n); break; }      // gcd = u = v
        if (c > 0) { bn_sub(u, v, n); bn_sub_mod(x1, x2, m, n); }
        else { bn_sub(v, u, n); bn_sub_mod(x2, x1, m, n); }
    }
    uint32_t* out = P.out + (size_t)i * P.limbs;
    const uint32_t* res = bn_is_one(u, n) ? x1 : x2;
    for (int k = 0; k < n; ++k) out[k] = ok ? res[k] : 0;
    if (!ok) atomicMin(P.first_bad, i);
}


// ---------------------------------------------------------------- warp-cooperative modinv
// The same binary extended Euclid with one WARP per item: lane l holds limbs [l*W, (l+1)*W).  Shifts take the crossing
// bit from the neighbour lane, additions / subtractions resolve the cross-lane carries with a ballot look-ahead, and a
// comparison is one pair of ballots (the highest differing lane decides).  ~50x lower latency than the one-thread
// kernel, which matters because Montgomery's batch inversion funnels a whole batch into one such inversion per chunk.
template <int W>
struct WarpNum {
    uint32_t v[W];
};

template <int W>
__device__ __forceinline__ bool wn_is_one(const WarpNum<W>& a, int lane) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < W; ++k) ok = ok && a.v[k] == ((lane == 0 && k == 0) ? 1u : 0u);
    return __all_sync(0xffffffffu, ok);
}
template <int W>
__device__ __forceinline__ bool wn_is_zero(const WarpNum<W>& a) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < W; ++k) ok = ok && a.v[k] == 0;
    return __all_sync(0xffffffffu, ok);
}
template <int W>
__device__ __forceinline__ bool wn_is_even(const WarpNum<W>& a) {
    return (__shfl_sync(0xffffffffu, a.v[0], 0) & 1u) == 0;
}
// a >>= 1 with `top` shifted into the most significant bit
template <int W>
__device__ __forceinline__ void wn_shr1(WarpNum<W>& a, uint32_t top, int lane) {
    uint32_t next = __shfl_down_sync(0xffffffffu, a.v[0], 1);
    if (lane == 31) next = top;
#pragma unroll
    for (int k = 0; k < W - 1; ++k) a.v[k] = (a.v[k] >> 1) | (a.v[k + 1] << 31);
    a.v[W - 1] = (a.v[W - 1] >> 1) | (next << 31);
}
// a += b, returns the carry out of the whole number
template <int W>
__device__ __forceinline__ uint32_t wn_add(WarpNum<W>& a, const WarpNum<W>& b, int lane) {
    uint64_t c = 0;
    uint32_t all1 = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < W; ++k) { c += (uint64_t)a.v[k] + b.v[k]; a.v[k] = (uint32_t)c; c >>= 32; all1 &= a.v[k]; }
    const uint32_t g = __ballot_sync(0xffffffffu, c != 0);
    const uint32_t p = __ballot_sync(0xffffffffu, all1 == 0xffffffffu) & ~g;
    const uint64_t ci = ((uint64_t)(g | p) + g) ^ p;          // bit l: carry into lane l; bit 32: carry out
    uint32_t cin = (uint32_t)(ci >> lane) & 1u;
#pragma unroll
    for (int k = 0; k < W; ++k) { const uint32_t t = a.v[k] + cin; cin = (t < cin) ? 1u : 0u; a.v[k] = t; }
    return (uint32_t)(ci >> 32) & 1u;
}
// a -= b, returns the borrow out of the whole number
template <int W>
__device__ __forceinline__ uint32_t wn_sub(WarpNum<W>& a, const WarpNum<W>& b, int lane) {
    int64_t bw = 0;
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) { const int64_t d = (int64_t)a.v[k] - b.v[k] - bw; bw = d < 0; a.v[k] = (uint32_t)d; any |= a.v[k]; }
    const uint32_t g = __ballot_sync(0xffffffffu, bw != 0);
    const uint32_t p = __ballot_sync(0xffffffffu, any == 0) & ~g;
    const uint64_t bi = ((uint64_t)(g | p) + g) ^ p;
    uint32_t bin = (uint32_t)(bi >> lane) & 1u;
#pragma unroll
    for (int k = 0; k < W; ++k) { const uint32_t t = a.v[k] - bin; bin = (a.v[k] < bin) ? 1u : 0u; a.v[k] = t; }
    return (uint32_t)(bi >> 32) & 1u;
}
// -1, 0, 1
template <int W>
__device__ __forceinline__ int wn_cmp(const WarpNum<W>& a, const WarpNum<W>& b) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) if (a.v[k] != b.v[k]) c = a.v[k] < b.v[k] ? -1 : 1;      // the highest differing limb wins
    const uint32_t gt = __ballot_sync(0xffffffffu, c > 0), lt = __ballot_sync(0xffffffffu, c < 0);
    return gt == lt ? 0 : (gt > lt ? 1 : -1);
}
template <int W>
__device__ __forceinline__ void wn_half_mod(WarpNum<W>& x, const WarpNum<W>& m, int lane) {
    uint32_t top = 0;
    if (!wn_is_even(x)) top = wn_add(x, m, lane);
    wn_shr1(x, top, lane);
}
template <int W>
__device__ __forceinline__ void wn_sub_mod(WarpNum<W>& x, const WarpNum<W>& y, const WarpNum<W>& m, int lane) {
    if (wn_sub(x, y, lane)) wn_add(x, m, lane);
}

template <int W>
__global__ void __launch_bounds__(128) modinv_warp_kernel(InvParams P) {
    const int lane = threadIdx.x & 31;
    const uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= P.n_items) return;                        // whole warps leave together
    WarpNum<W> u, v, x1, x2, m;
    const uint32_t* a = P.in + (size_t)item * P.limbs + lane * W;
#pragma unroll
    for (int k = 0; k < W; ++k) { m.v[k] = P.mod[lane * W + k]; u.v[k] = a[k]; v.v[k] = m.v[k]; x1.v[k] = 0; x2.v[k] = 0; }
    if (lane == 0) x1.v[0] = 1;
    while (wn_cmp(u, m) >= 0) wn_sub(u, m, lane);
    bool ok = !wn_is_zero(u);
    while (ok && !wn_is_one(u, lane) && !wn_is_one(v, lane)) {
        while (wn_is_even(u)) { wn_shr1(u, 0, lane); wn_half_mod(x1, m, lane); }
        while (wn_is_even(v)) { wn_shr1(v, 0, lane); wn_half_mod(x2, m, lane); }
        const int c = wn_cmp(u, v);
        if (c == 0) { ok = wn_is_one(u, lane); break; }
        if (c > 0) { wn_sub(u, v, lane); wn_sub_mod(x1, x2, m, lane); }
        else { wn_sub(v, u, lane); wn_sub_mod(x2, x1, m, lane); }
    }
    const bool use1 = wn_is_one(u, lane);
    uint32_t* out = P.out + (size_t)item * P.limbs + lane * W;
#pragma unroll
    for (int k = 0; k < W; ++k) out[k] = ok ? (use1 ? x1.v[k] : x2.v[k]) : 0u;
    if (!ok && lane == 0) atomicMin(P.first_bad, item);
}

template <int W>
static cudaError_t modinv_warp_launch(const InvParams& P, cudaStream_t stream) {
    const unsigned blocks = (P.n_items + 3) / 4;          // 4 warps = 4 items per block
    modinv_warp_kernel<W><<<blocks, 128, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t modinv_launch(const InvParams& P, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    switch (P.limbs) {
        case 32: return modinv_warp_launch<1>(P, stream);
        case 64: return modinv_warp_launch<2>(P, stream);
        case 96: return modinv_warp_launch<3>(P, stream);
        case 128: return modinv_warp_launch<4>(P, stream);
        case 192: return modinv_warp_launch<6>(P, stream);
        default: break;
    }
    const int threads = 32;
    modinv_kernel<<<(P.n_items + threads - 1) / threads, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

// ------------------------------------------------------- unreduced products
// out (na+nb limbs) = a * b, product scanning with a 96-bit column accumulator
__global__ void bigmul_kernel(MulParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_items) return;
    const uint32_t* a = P.a + (size_t)i * P.a_stride;
    const uint32_t* b = P.b + (size_t)i * P.b_stride;
    uint32_t* out = P.out + (size_t)i * P.out_stride;
    const int na = P.na, nb = P.nb;
    uint64_t acc = 0; uint32_t acc_hi = 0;
    for (int k = 0; k < na + nb - 1; ++k) {
        const int lo = k - (nb - 1) > 0 ? k - (nb - 1) : 0, hi = k < na - 1 ? k : na - 1;
        for (int ia = lo; ia <= hi; ++ia) {
            const uint64_t p = (uint64_t)a[ia] * b[k - ia];
            acc += p;
            acc_hi += acc < p;
        }
        out[k] = (uint32_t)acc;
        acc = (acc >> 32) | ((uint64_t)acc_hi << 32);
        acc_hi = 0;
    }
    out[na + nb - 1] = (uint32_t)acc;
    for (uint32_t k = na + nb; k < P.out_limbs; ++k) out[k] = 0;
}

cudaError_t bigmul_launch(const MulParams& P, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    const int threads = 64;
    bigmul_kernel<<<(P.n_items + threads - 1) / threads, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

// out (out_limbs) = r + e * k, k a per-key constant (Z = r + E*delta*share)
__global__ void muladd_kernel(MulAddParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_items) return;
    const uint32_t* r = P.r + (size_t)i * P.r_stride;
    const uint32_t* e = P.e + (size_t)i * P.e_stride;
    uint32_t* out = P.out + (size_t)i * P.out_stride;
    const int ne = P.ne, nk = P.nk, nout = P.out_limbs;
    for (int k = 0; k < nout; ++k) out[k] = k < (int)P.nr ? r[k] : 0;
    for (int ie = 0; ie < ne; ++ie) {
        uint64_t c = 0;
        const uint32_t ei = e[ie];
        for (int j = 0; j < nk && ie + j < nout; ++j) {
            c += (uint64_t)ei * P.k[j] + out[ie + j];
            out[ie + j] = (uint32_t)c; c >>= 32;
        }
        for (int j = ie + nk; c != 0 && j < nout; ++j) { c += out[j]; out[j] = (uint32_t)c; c >>= 32; }
    }
}

cudaError_t muladd_launch(const MulAddParams& P, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    const int threads = 64;
    muladd_kernel<<<(P.n_items + threads - 1) / threads, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ SHA-256
__constant__ uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

struct Sha256 {
    uint32_t h[8];
    uint32_t w[16];      // current block, big-endian words
    uint32_t fill;       // bytes in the current block
    uint64_t total;      // bytes hashed so far
    __device__ void init() {
        h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
        h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
        fill = 0; total = 0;
        for (int i = 0; i < 16; ++i) w[i] = 0;
    }
    __device__ static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    __device__ void compress() {
        uint32_t m[64];
        for (int i = 0; i < 16; ++i) m[i] = w[i];
        for (int i = 16; i < 64; ++i) {
            const uint32_t s0 = rotr(m[i - 15], 7) ^ rotr(m[i - 15], 18) ^ (m[i - 15] >> 3);
            const uint32_t s1 = rotr(m[i - 2], 17) ^ rotr(m[i - 2], 19) ^ (m[i - 2] >> 10);
            m[i] = m[i - 16] + s0 + m[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t t1 = hh + S1 + ch + SHA_K[i] + m[i];
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
        for (int i = 0; i < 16; ++i) w[i] = 0;
        fill = 0;
    }
    __device__ void put(uint8_t byte) {
        w[fill >> 2] |= (uint32_t)byte << (24 - 8 * (fill & 3));
        ++fill; ++total;
        if (fill == 64) compress();
    }
    // minimal big-endian magnitude of a little-endian limb array (gmp.Int.Bytes(); zero -> no bytes)
    __device__ void put_int(const uint32_t* a, int n) {
        int top = n - 1;
        while (top >= 0 && a[top] == 0) --top;
        if (top < 0) return;
        const uint32_t t = a[top];
        const int nb = t >> 24 ? 4 : t >> 16 ? 3 : t >> 8 ? 2 : 1;
        for (int b = nb - 1; b >= 0; --b) put((uint8_t)(t >> (8 * b)));
        for (int k = top - 1; k >= 0; --k) {
            const uint32_t x = a[k];
            if ((fill & 3) == 0 && fill <= 60) {   // aligned fast path
                w[fill >> 2] = x; fill += 4; total += 4;
                if (fill == 64) compress();
            } else {
                put((uint8_t)(x >> 24)); put((uint8_t)(x >> 16)); put((uint8_t)(x >> 8)); put((uint8_t)x);
            }
        }
    }
    __device__ void finish() {
        const uint64_t bits = total * 8;
        put(0x80);
        total -= 1;
        if (fill > 56) compress();
        w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits;
        compress();
    }
};

// digest[i] = SHA-256(Bytes(seg0[i]) || Bytes(seg1[i]) || ...), written as the integer
// SetBytes(digest) in 8 little-endian limbs (E of the ZKP; the low bit is the DDLEQ challenge)
__global__ void sha256_concat_kernel(ShaParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_items) return;
    Sha256 S; S.init();
    for (int s = 0; s < P.n_seg; ++s) S.put_int(P.seg[s] + (size_t)(i / (P.div[s] ? P.div[s] : 1u)) * P.stride[s], P.limbs[s]);
    S.finish();
    uint32_t* out = P.out + (size_t)i * 8;
    for (int k = 0; k < 8; ++k) out[k] = S.h[7 - k];
}

cudaError_t sha256_concat_launch(const ShaParams& P, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    const int threads = 64;
    sha256_concat_kernel<<<(P.n_items + threads - 1) / threads, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

// flags[i] = (a[i] == b[i]) over records of `limbs` limbs
__global__ void equal_kernel(const uint32_t* a, const uint32_t* b, uint32_t limbs, uint32_t n_items, uint8_t* flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    bool eq = true;
    for (uint32_t k = 0; k < limbs; ++k) eq = eq && a[(size_t)i * limbs + k] == b[(size_t)i * limbs + k];
    flags[i] = eq ? 1 : 0;
}

cudaError_t equal_launch(const uint32_t* a, const uint32_t* b, uint32_t limbs, uint32_t n_items, uint8_t* flags, cudaStream_t stream) {
    if (n_items == 0) return cudaSuccess;
    equal_kernel<<<(n_items + 127) / 128, 128, 0, stream>>>(a, b, limbs, n_items, flags);
    return cudaGetLastError();
}

__global__ void first_zero_kernel(const uint8_t* flags, uint32_t n_items, uint32_t* first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && blockIdx.x == 0) atomicMin(first, 0xffffffffu);
    if (i < n_items && flags[i] == 0) atomicMin(first, i);
}

cudaError_t first_zero_launch(const uint8_t* flags, uint32_t n_items, uint32_t* first, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(first, 0xff, 4, stream);
    if (e != cudaSuccess || n_items == 0) return e;
    first_zero_kernel<<<(n_items + 127) / 128, 128, 0, stream>>>(flags, n_items, first);
    return cudaGetLastError();
}

// out[i] = pick[i] ? a[i] : b[i]  (records of `limbs` limbs; pick = low bit of an 8-limb digest record)
__global__ void select_kernel(const uint32_t* digest, const uint32_t* a, uint32_t a_div, const uint32_t* b, uint32_t b_div, uint32_t limbs, uint32_t n_items, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const uint32_t* src = (digest[(size_t)i * 8] & 1) ? a + (size_t)(i / a_div) * limbs : b + (size_t)(i / b_div) * limbs;
    for (uint32_t k = 0; k < limbs; ++k) out[(size_t)i * limbs + k] = src[k];
}

cudaError_t select_launch(const uint32_t* digest, const uint32_t* a, uint32_t a_div, const uint32_t* b, uint32_t b_div, uint32_t limbs, uint32_t n_items, uint32_t* out, cudaStream_t stream) {
    if (n_items == 0) return cudaSuccess;
    select_kernel<<<(n_items + 127) / 128, 128, 0, stream>>>(digest, a, a_div ? a_div : 1, b, b_div ? b_div : 1, limbs, n_items, out);
    return cudaGetLastError();
}

__global__ void repeat_kernel(const uint32_t* in, uint32_t limbs, uint32_t rep, uint32_t* out, uint32_t n_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const uint32_t* src = in + (size_t)(i / rep) * limbs;
    for (uint32_t k = 0; k < limbs; ++k) out[(size_t)i * limbs + k] = src[k];
}

cudaError_t repeat_launch(const uint32_t* in, uint32_t limbs, uint32_t rep, uint32_t* out, uint32_t n_out, cudaStream_t stream) {
    if (n_out == 0) return cudaSuccess;
    repeat_kernel<<<(n_out + 127) / 128, 128, 0, stream>>>(in, limbs, rep, out, n_out);
    return cudaGetLastError();
}

__global__ void gather_kernel(const uint32_t* in, uint32_t limbs, const uint32_t* idx, uint32_t div, uint32_t* out, uint32_t n) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint32_t* src = in + (size_t)(idx[i] / div) * limbs;
    for (uint32_t k = lane; k < limbs; k += 32) out[(size_t)i * limbs + k] = src[k];
}
__global__ void scatter_kernel(const uint32_t* in, uint32_t limbs, const uint32_t* idx, uint32_t* out, uint32_t n) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (i >= n) return;
    uint32_t* dst = out + (size_t)idx[i] * limbs;
    for (uint32_t k = lane; k < limbs; k += 32) dst[k] = in[(size_t)i * limbs + k];
}
cudaError_t gather_launch(const uint32_t* in, uint32_t limbs, const uint32_t* idx, uint32_t div, uint32_t* out, uint32_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    gather_kernel<<<(n + 3) / 4, 128, 0, stream>>>(in, limbs, idx, div ? div : 1, out, n);
    return cudaGetLastError();
}
cudaError_t scatter_launch(const uint32_t* in, uint32_t limbs, const uint32_t* idx, uint32_t* out, uint32_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    scatter_kernel<<<(n + 3) / 4, 128, 0, stream>>>(in, limbs, idx, out, n);
    return cudaGetLastError();
}

// widen / narrow records: out (out_limbs) = in (in_limbs), zero padded or truncated; stride 0 broadcasts one record
__global__ void resize_kernel(const uint32_t* in, uint32_t in_stride, uint32_t in_limbs, uint32_t* out, uint32_t out_limbs, uint32_t n_items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    for (uint32_t k = 0; k < out_limbs; ++k) out[(size_t)i * out_limbs + k] = k < in_limbs ? in[(size_t)i * in_stride + k] : 0;
}

cudaError_t resize_launch(const uint32_t* in, uint32_t in_stride, uint32_t in_limbs, uint32_t* out, uint32_t out_limbs, uint32_t n_items, cudaStream_t stream) {
    if (n_items == 0) return cudaSuccess;
    resize_kernel<<<(n_items + 127) / 128, 128, 0, stream>>>(in, in_stride, in_limbs, out, out_limbs, n_items);
    return cudaGetLastError();
}

}  // namespace pgpu
