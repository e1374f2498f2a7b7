// Safe-prime candidate testing (safe_prime.go:147-290) for a batch of candidates:
//   sieve_kernel  one thread per candidate: byte masking, residue mod 3*5*...*53, the delta search
//                 with the reference's cumulative q += delta, q = 1 (mod 3) and p = 2q+1 sieves;
//   strong_kernel warp-cooperative Montgomery arithmetic with a PER-ITEM modulus: Miller-Rabin
//                 strong test to a small base (q.ProbablyPrime) or the base-2 Fermat test
//                 2^(p-1) = 1 (mod p) of Pocklington's criterion (safe_prime.go:272-278).
// Base 2 needs no multiplications besides the squarings: "times 2" is a modular doubling.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/pgpu.h"
#include "mont.cuh"

namespace pgpu {

namespace {

__constant__ uint32_t SMALL_PRIMES[15] = {3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53};   // safe_prime.go:26-28
constexpr uint64_t SMALL_PRIMES_PRODUCT = 16294579238595022365ull;                                    // safe_prime.go:34

struct SieveParams {
    uint32_t n_items;
    uint32_t q_bits;         // qBitLen = pBitLen - 1
    uint32_t raw_bytes;      // (qBitLen + 7) / 8
    uint32_t S;              // limbs per output record
    const uint8_t* raw;
    uint32_t* q;             // q0 + cumulative delta
    uint32_t* p;             // 2q + 1 (0 if no delta survived)
    uint8_t* state;          // 1 = test it, 0 = rejected without a primality test (bit length changed / no delta)
};

__device__ __forceinline__ bool hits_small_prime(uint64_t m, bool allow_equal) {
#pragma unroll
    for (int i = 0; i < 15; ++i) {
        const uint64_t pr = SMALL_PRIMES[i];
        if (m % pr == 0 && !(allow_equal && m == pr)) return true;
    }
    return false;
}

__global__ void sieve_kernel(SieveParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_items) return;
    const uint8_t* raw = P.raw + (size_t)i * P.raw_bytes;
    uint32_t* q = P.q + (size_t)i * P.S;
    uint32_t* p = P.p + (size_t)i * P.S;
    const uint32_t nb = P.raw_bytes;
    uint32_t b = P.q_bits % 8; if (b == 0) b = 8;
    // big-endian bytes -> little-endian limbs with the masks of safe_prime.go:175-202
    for (uint32_t k = 0; k < P.S; ++k) q[k] = 0;
    for (uint32_t j = 0; j < nb; ++j) {
        uint32_t byte = raw[j];
        if (j == 0) {
            byte &= (1u << b) - 1u;
            if (b >= 2) byte |= 3u << (b - 2); else byte |= 1u;
        }
        if (j == 1 && b < 2) byte |= 0x80u;
        if (j == nb - 1) byte |= 1u;
        const uint32_t pos = nb - 1 - j;
        q[pos >> 2] |= byte << (8 * (pos & 3));
    }
    // mod = q mod smallPrimesProduct (:208-209)
    uint64_t mod = 0;
    for (int k = (int)P.S - 1; k >= 0; --k) mod = (uint64_t)((((unsigned __int128)mod << 32) | q[k]) % SMALL_PRIMES_PRODUCT);
    // NextDelta loop (:211-249).  q is advanced cumulatively (:221-224) while m keeps using the original mod.
    const bool small = P.q_bits <= 6;
    uint64_t off = 0;
    bool found = false;
    for (uint64_t delta = 0; delta < (1ull << 20); delta += 2) {
        const uint64_t m = mod + delta;
        if (hits_small_prime(m, small)) continue;
        off += delta;
        const uint64_t qm = (uint64_t)(((unsigned __int128)mod + off) % SMALL_PRIMES_PRODUCT);   // actual q mod P
        if (qm % 3 == 1) continue;                                                                // :238-241
        const uint64_t pm = (uint64_t)(((unsigned __int128)qm * 2 + 1) % SMALL_PRIMES_PRODUCT);   // p mod P, :244-247
        if (hits_small_prime(pm, true)) continue;                                                 // isPrimeCandidate(p), m != prime always allowed (:283-286)
        found = true;
        break;
    }
    // q += off
    uint64_t c = off;
    for (uint32_t k = 0; k < P.S && c; ++k) { c += q[k]; q[k] = (uint32_t)c; c >>= 32; }
    // bit length check (:256-258) and p = 2q + 1
    int top = (int)P.S - 1;
    while (top > 0 && q[top] == 0) --top;
    const uint32_t bits = 32u * top + (32u - __clz(q[top]));
    bool ok = found && bits == P.q_bits && c == 0;
    uint32_t carry = 1;
    for (uint32_t k = 0; k < P.S; ++k) {
        const uint32_t v = q[k];
        p[k] = found ? ((v << 1) | carry) : 0u;
        carry = v >> 31;
    }
    if (carry) ok = false;        // p does not fit the record (cannot happen when the bit length is right)
    P.state[i] = ok ? 1 : 0;
}

struct StrongParams {
    uint32_t n_items;            // entries of idx (or items if idx == nullptr)
    const uint32_t* idx;         // items to test, or nullptr for 0..n_items-1
    const uint32_t* mod;         // records of S limbs, odd
    uint32_t bits;               // bit length of every modulus
    uint32_t base;               // small base a >= 2 (used when bases == nullptr)
    const uint32_t* bases;       // nb bases: slot s tests item idx[s / nb] to base bases[s % nb]
    uint32_t nb;
    uint32_t fermat;             // 0: Miller-Rabin strong test; 1: Fermat test a^(m-1) = 1
    uint8_t* flags;              // flags[item] &= pass
};

template <int TPI, int L>
__global__ void __launch_bounds__(128) strong_kernel(StrongParams P) {
    constexpr int S = TPI * L;
    Mont<TPI, L> M;
    const uint32_t n_groups = gridDim.x * blockDim.x / TPI;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / TPI;
    const uint32_t rounds = (P.n_items + n_groups - 1) / n_groups;
    const int lane_t = (threadIdx.x & 31) & (TPI - 1);
    for (uint32_t rd = 0; rd < rounds; ++rd) {
        uint32_t slot = rd * n_groups + group;
        const bool active = slot < P.n_items;
        if (!active) slot = 0;
        const uint32_t nb = P.bases ? P.nb : 1u;
        const uint32_t which = slot / nb;
        const uint32_t item = P.idx ? P.idx[which] : which;
        const uint32_t base = P.bases ? P.bases[slot - which * nb] : P.base;
        const uint32_t* mod = P.mod + (size_t)item * S;
        // -m^-1 mod 2^32 by Newton iteration on the lowest limb
        const uint32_t m0 = mod[0];
        uint32_t inv = m0;
#pragma unroll
        for (int it = 0; it < 5; ++it) inv *= 2u - m0 * inv;
        M.init(mod, 0u - inv);
        // one = R mod m: (2^bits - m) doubled (32*S - bits) times
        uint32_t one[L];
#pragma unroll
        for (int k = 0; k < L; ++k) one[k] = ~M.n[k];
        if (lane_t == 0) {
            uint32_t cy = 1;
#pragma unroll
            for (int k = 0; k < L; ++k) { const uint32_t v = one[k] + cy; cy = (v < cy) ? 1u : 0u; one[k] = v; }
        }
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const uint32_t pos = 32u * (lane_t * L + k);
            if (pos >= P.bits) one[k] = 0;
            else if (pos + 32 > P.bits) one[k] &= (1u << (P.bits - pos)) - 1u;
        }
        for (uint32_t d = P.bits; d < 32u * S; ++d) M.add(one, one, one);
        // minus_one = m - one; aR = a * one by double-and-add
        uint32_t minus1[L], zero[L], aR[L];
#pragma unroll
        for (int k = 0; k < L; ++k) zero[k] = 0;
        M.sub(minus1, zero, one);
        // aR = a * one by double-and-add over a fixed number of bit positions: groups of one warp may hold different bases
        // and every Mont operation is a warp-wide collective, so the trip count must not depend on the base (a < 2^8)
#pragma unroll
        for (int k = 0; k < L; ++k) aR[k] = 0;
        for (int bpos = 7; bpos >= 0; --bpos) {
            M.add(aR, aR, aR);
            uint32_t t[L];
            M.add(t, aR, one);
            const bool take = (base >> bpos) & 1u;
#pragma unroll
            for (int k = 0; k < L; ++k) aR[k] = take ? t[k] : aR[k];
        }
        // exponent e = m - 1 (m odd: clear bit 0); s = trailing zeros of e (group-uniform)
        uint32_t e[L];
#pragma unroll
        for (int k = 0; k < L; ++k) e[k] = M.n[k];
        if (lane_t == 0) e[0] &= ~1u;
        uint32_t tz = 32u * L;      // trailing zeros inside this lane's limbs
#pragma unroll
        for (int k = L - 1; k >= 0; --k) if (e[k]) tz = 32u * k + (__ffs(e[k]) - 1);
        uint32_t s = 32u * S;
        for (int t = TPI - 1; t >= 0; --t) {
            const uint32_t z = __shfl_sync(FULL_MASK, tz, t, TPI);
            if (z < 32u * L) s = 32u * L * t + z;
        }
        uint32_t x[L];
#pragma unroll
        for (int k = 0; k < L; ++k) x[k] = aR[k];
        bool pass = false;
        const bool base2 = P.bases == nullptr && P.base == 2;
        const int last = P.fermat ? 0 : 1;
        for (int i = (int)P.bits - 2; i >= last; --i) {
            M.mul(x, x, x);
            // bit i of e, broadcast from the owning lane
            const int limb = i >> 5, owner = limb / L;
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < L; ++k) if (k == limb - owner * L) w = e[k];
            w = __shfl_sync(FULL_MASK, w, owner, TPI);
            const bool bit = (w >> (i & 31)) & 1u;
            uint32_t y[L];
            if (base2) {
                M.add(y, x, x);
            } else {
                uint32_t mulby[L];
#pragma unroll
                for (int k = 0; k < L; ++k) mulby[k] = bit ? aR[k] : one[k];
                M.mul(y, x, mulby);
            }
#pragma unroll
            for (int k = 0; k < L; ++k) x[k] = (bit || !base2) ? y[k] : x[k];
            if (!P.fermat) {
                bool eq1 = true, eqm = true;
#pragma unroll
                for (int k = 0; k < L; ++k) { eq1 = eq1 && x[k] == one[k]; eqm = eqm && x[k] == minus1[k]; }
                const bool all1 = M.gballot(eq1) == Mont<TPI, L>::GMASK, allm = M.gballot(eqm) == Mont<TPI, L>::GMASK;
                if ((uint32_t)i == s && (all1 || allm)) pass = true;
                if ((uint32_t)i < s && allm) pass = true;
            }
        }
        if (P.fermat) {
            bool eq1 = true;
#pragma unroll
            for (int k = 0; k < L; ++k) eq1 = eq1 && x[k] == one[k];
            pass = M.gballot(eq1) == Mont<TPI, L>::GMASK;
        }
        if (active && lane_t == 0 && !pass) P.flags[item] = 0;
    }
}

template <int TPI, int L>
cudaError_t strong_launch_t(const StrongParams& P, int sms, cudaStream_t stream) {
    if (P.n_items == 0) return cudaSuccess;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, strong_kernel<TPI, L>, 128, 0);
    if (per_sm <= 0) return cudaErrorInvalidValue;
    const int gpb = 128 / TPI;
    const int want = (int)((P.n_items + gpb - 1) / gpb);
    const int blocks = std::min(want, sms * per_sm);
    strong_kernel<TPI, L><<<blocks, 128, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t strong_launch(int S, const StrongParams& P, int sms, cudaStream_t stream) {
    switch (S) {
        case 32: return strong_launch_t<4, 8>(P, sms, stream);
        case 48: return strong_launch_t<4, 12>(P, sms, stream);
        case 64: return strong_launch_t<4, 16>(P, sms, stream);
    }
    return cudaErrorInvalidValue;
}

thread_local std::string g_perr;
int pfail(int code, const std::string& m) { g_perr = m; return code; }

#define PCU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return pfail(PGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)

const uint32_t MR_BASES[20] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71};

// one stream per device, created once; scratch comes from the stream-ordered pool (kept warm across calls)
cudaError_t scratch_stream(int device, cudaStream_t* st) {
    static cudaStream_t streams[64] = {};
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (!streams[device]) {
        cudaError_t e = cudaStreamCreateWithFlags(&streams[device], cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    *st = streams[device];
    return cudaSuccess;
}

int shape_for_bits(unsigned bits) { return bits <= 1024 ? 32 : bits <= 1536 ? 48 : bits <= 2048 ? 64 : 0; }

}  // namespace

const char* primes_last_error() { return g_perr.c_str(); }

// flags[i] = 1 if cand[i] (records of S limbs, all of `bits` bits, odd) passes `rounds` Miller-Rabin rounds
// (bases 2, 3, 5, ...).  Items that fail base 2 are not tested further.
static int mr_filter(int S, int sms, cudaStream_t st, size_t count, const uint32_t* d_cand, uint32_t bits, unsigned rounds,
                     uint8_t* d_flags, std::vector<uint8_t>& h_flags, uint32_t* d_idx, const uint32_t* d_bases, uint64_t* launches, std::string& err) {
    StrongParams P{(uint32_t)count, nullptr, d_cand, bits, 2, nullptr, 0, 0, d_flags};
    cudaError_t e = strong_launch(S, P, sms, st);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return PGPU_ERR_CUDA; }
    ++*launches;
    h_flags.resize(count);
    e = cudaMemcpyAsync(h_flags.data(), d_flags, count, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return PGPU_ERR_CUDA; }
    std::vector<uint32_t> idx;
    for (size_t i = 0; i < count; ++i) if (h_flags[i]) idx.push_back((uint32_t)i);
    if (idx.empty() || rounds <= 1) return PGPU_OK;
    e = cudaMemcpyAsync(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return PGPU_ERR_CUDA; }
    {   // the remaining bases of every survivor in one launch
        const unsigned nb = std::min(rounds, 20u) - 1;
        StrongParams Q{(uint32_t)(idx.size() * nb), d_idx, d_cand, bits, 0, d_bases + 1, nb, 0, d_flags};
        e = strong_launch(S, Q, sms, st);
        if (e != cudaSuccess) { err = cudaGetErrorString(e); return PGPU_ERR_CUDA; }
        ++*launches;
    }
    e = cudaMemcpyAsync(h_flags.data(), d_flags, count, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return PGPU_ERR_CUDA; }
    return PGPU_OK;
}

}  // namespace pgpu

using namespace pgpu;

#pragma GCC visibility push(default)
extern "C" {

const char* pgpu_primes_last_error(void) { return primes_last_error(); }

int pgpu_miller_rabin(int device, unsigned bits, size_t count, const void* cand, unsigned rounds, uint8_t* ok, uint64_t* launches) {
    const int S = shape_for_bits(bits);
    if (!S || bits < 8) return pfail(PGPU_ERR_UNSUPPORTED, "pgpu_miller_rabin: candidates of 8..2048 bits are supported");
    if (count && (!cand || !ok)) return pfail(PGPU_ERR_ARG, "pgpu_miller_rabin: null argument");
    if (rounds < 1 || rounds > 20) return pfail(PGPU_ERR_ARG, "pgpu_miller_rabin: 1..20 rounds");
    if (count == 0) return PGPU_OK;
    for (size_t i = 0; i < count; ++i) {      // the kernel derives R mod m from the common bit length
        const uint32_t* c = (const uint32_t*)cand + i * S;
        unsigned top = S; while (top > 0 && c[top - 1] == 0) --top;
        const unsigned b = top ? 32 * (top - 1) + (32 - __builtin_clz(c[top - 1])) : 0;
        if (b != bits || !(c[0] & 1)) return pfail(PGPU_ERR_ARG, "pgpu_miller_rabin: candidate " + std::to_string(i) + " is even or not of the stated bit length");
    }
    uint32_t *d_cand = nullptr, *d_idx = nullptr, *d_bases = nullptr; uint8_t* d_flags = nullptr; cudaStream_t st = nullptr;
    auto cleanup = [&]() { for (void* x : {(void*)d_cand, (void*)d_idx, (void*)d_bases, (void*)d_flags}) if (x) cudaFreeAsync(x, st); if (st) cudaStreamSynchronize(st); };
    PCU(cudaSetDevice(device));
    int sms_attr = 0; PCU(cudaDeviceGetAttribute(&sms_attr, cudaDevAttrMultiProcessorCount, device));
    PCU(scratch_stream(device, &st));
    PCU(cudaMallocAsync(&d_cand, count * S * 4, st)); PCU(cudaMallocAsync(&d_idx, count * 4, st)); PCU(cudaMallocAsync(&d_flags, count, st)); PCU(cudaMallocAsync(&d_bases, sizeof MR_BASES, st));
    PCU(cudaMemcpyAsync(d_bases, MR_BASES, sizeof MR_BASES, cudaMemcpyHostToDevice, st));
    PCU(cudaMemcpyAsync(d_cand, cand, count * S * 4, cudaMemcpyHostToDevice, st));
    PCU(cudaMemsetAsync(d_flags, 1, count, st));
    std::vector<uint8_t> h; std::string err; uint64_t nl = 0;
    int rc = mr_filter(S, sms_attr, st, count, d_cand, bits, rounds, d_flags, h, d_idx, d_bases, &nl, err);
    if (rc == PGPU_OK) memcpy(ok, h.data(), count);
    if (launches) *launches = nl;
    cleanup();
    return rc ? pfail(rc, err) : PGPU_OK;
}

int pgpu_safe_prime_scan(int device, unsigned p_bits, size_t count, const uint8_t* raw, void* p_out, void* q_out, uint8_t* ok,
                         uint64_t* launches) {
    const int S = shape_for_bits(p_bits);
    if (!S || p_bits < 8) return pfail(PGPU_ERR_UNSUPPORTED, "pgpu_safe_prime_scan: p of 8..2048 bits is supported");
    if (count && (!raw || !p_out || !q_out || !ok)) return pfail(PGPU_ERR_ARG, "pgpu_safe_prime_scan: null argument");
    if (count == 0) return PGPU_OK;
    if (count > 0x7fffffffu) return pfail(PGPU_ERR_ARG, "pgpu_safe_prime_scan: batch too large");
    const unsigned q_bits = p_bits - 1, raw_bytes = (q_bits + 7) / 8;
    uint8_t *d_raw = nullptr, *d_state = nullptr, *d_pf = nullptr; uint32_t *d_q = nullptr, *d_p = nullptr, *d_idx = nullptr, *d_bases = nullptr; cudaStream_t st = nullptr;
    auto cleanup = [&]() {
        for (void* x : {(void*)d_raw, (void*)d_state, (void*)d_pf, (void*)d_q, (void*)d_p, (void*)d_idx, (void*)d_bases}) if (x) cudaFreeAsync(x, st);
        if (st) cudaStreamSynchronize(st);
    };
    PCU(cudaSetDevice(device));
    int sms = 0; PCU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    PCU(scratch_stream(device, &st));
    PCU(cudaMallocAsync(&d_raw, count * raw_bytes, st)); PCU(cudaMallocAsync(&d_state, count, st)); PCU(cudaMallocAsync(&d_pf, count, st));
    PCU(cudaMallocAsync(&d_q, count * S * 4, st)); PCU(cudaMallocAsync(&d_p, count * S * 4, st)); PCU(cudaMallocAsync(&d_idx, count * 4, st)); PCU(cudaMallocAsync(&d_bases, sizeof MR_BASES, st));
    PCU(cudaMemcpyAsync(d_bases, MR_BASES, sizeof MR_BASES, cudaMemcpyHostToDevice, st));
    PCU(cudaMemcpyAsync(d_raw, raw, count * raw_bytes, cudaMemcpyHostToDevice, st));
    uint64_t nl = 0;
    SieveParams SP{(uint32_t)count, q_bits, raw_bytes, (uint32_t)S, d_raw, d_q, d_p, d_state};
    sieve_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(SP);
    PCU(cudaGetLastError()); ++nl;
    // q.ProbablyPrime(20) (:256): base-2 strong test on everything the sieve let through, the other 19
    // bases only on its survivors
    std::vector<uint8_t> state(count);
    PCU(cudaMemcpyAsync(state.data(), d_state, count, cudaMemcpyDeviceToHost, st));
    PCU(cudaStreamSynchronize(st));
    std::vector<uint32_t> idx;
    for (size_t i = 0; i < count; ++i) if (state[i]) idx.push_back((uint32_t)i);
    std::vector<uint8_t> flags(count, 0);
    if (!idx.empty()) {
        PCU(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, st));
        StrongParams P{(uint32_t)idx.size(), d_idx, d_q, q_bits, 2, nullptr, 0, 0, d_state};
        PCU(strong_launch(S, P, sms, st)); ++nl;
        PCU(cudaMemcpyAsync(state.data(), d_state, count, cudaMemcpyDeviceToHost, st));
        PCU(cudaStreamSynchronize(st));
        idx.clear();
        for (size_t i = 0; i < count; ++i) if (state[i]) idx.push_back((uint32_t)i);
        if (!idx.empty()) {
            PCU(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, st));
            {   // bases 3 .. 71 of every survivor in one launch
                StrongParams Q{(uint32_t)(idx.size() * 19), d_idx, d_q, q_bits, 0, d_bases + 1, 19, 0, d_state};
                PCU(strong_launch(S, Q, sms, st)); ++nl;
            }
            // isPocklingtonCriterionSatisfied(p) (:257, :272-278): evaluated only when q passed (&& short circuit)
            StrongParams F{(uint32_t)idx.size(), d_idx, d_p, p_bits, 2, nullptr, 0, 1, d_state};
            PCU(strong_launch(S, F, sms, st)); ++nl;
            PCU(cudaMemcpyAsync(state.data(), d_state, count, cudaMemcpyDeviceToHost, st));
            PCU(cudaStreamSynchronize(st));
        }
        flags = state;
    }
    PCU(cudaMemcpyAsync(p_out, d_p, count * S * 4, cudaMemcpyDeviceToHost, st));
    PCU(cudaMemcpyAsync(q_out, d_q, count * S * 4, cudaMemcpyDeviceToHost, st));
    PCU(cudaStreamSynchronize(st));
    memcpy(ok, flags.data(), count);
    if (launches) *launches = nl;
    cleanup();
    return PGPU_OK;
}

}  // extern "C"
#pragma GCC visibility pop
