// Warp-cooperative, limb-sliced Montgomery arithmetic for sm_100a.
//
// One multi-precision residue of S = TPI*L 32-bit limbs is held by a group of
// TPI consecutive lanes of a warp; lane t of the group keeps limbs
// [t*L, t*L+L) in registers.  Montgomery multiplication is CIOS (coarsely
// integrated operand scanning): for every limb b_j of the second operand the
// group accumulates a*b_j, derives the quotient digit q from the lowest limb,
// accumulates n*q and shifts the accumulator down one limb.
//
// Per lane the accumulator is split into an "even" and an "odd" array of
// 64-bit columns so that every 32x32->64 product lands on an aligned register
// pair and is a single IMAD.WIDE.U32 with carry-in/carry-out (ptxas fuses
// each mad.lo.cc/madc.hi.cc pair).  The one-limb shift after each quotient
// digit is free: the arrays swap roles and the 64-bit re-alignment of the
// even array is folded into the addend operand of the next step's products.
// Carries that leave a lane's window are kept lazily in two small carry words
// and resolved across lanes only once per multiplication (ballot based
// carry look-ahead), followed by the conditional subtraction of n.
//
// This replaces what the reference does with one cgo call into libgmp's
// mpz_powm / mpz_mul+mpz_mod per big integer (e.g. /root/reference/paillier.go:213-216).
#pragma once
#include <cstdint>

namespace pgpu {

constexpr unsigned FULL_MASK = 0xffffffffu;

// ---- single-instruction carry-chain primitives (CC flag lives across them) --
__device__ __forceinline__ void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) {
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
}
__device__ __forceinline__ void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) {
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
}
__device__ __forceinline__ void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) {
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
}
__device__ __forceinline__ void add_cc(uint32_t& d, uint32_t a, uint32_t b) {
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void addc_cc(uint32_t& d, uint32_t a, uint32_t b) {
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void addc(uint32_t& d, uint32_t a, uint32_t b) {
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void sub_cc(uint32_t& d, uint32_t a, uint32_t b) {
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void subc_cc(uint32_t& d, uint32_t a, uint32_t b) {
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void subc(uint32_t& d, uint32_t a, uint32_t b) {
    asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}

// Carry look-ahead over a group: g = lanes that generate a carry, p = lanes
// that propagate one (mutually exclusive).  Returns the mask of lanes that
// receive a carry-in; bit TPI is the carry out of the whole group.
__device__ __forceinline__ uint64_t lookahead(uint32_t g, uint32_t p) {
    return ((uint64_t)(g | p) + g) ^ p;
}

template <int TPI, int L>
struct Mont {
    static_assert((L % 2) == 0, "L must be even");
    static_assert(TPI >= 1 && TPI <= 32 && (TPI & (TPI - 1)) == 0, "TPI must be a power of two");
    static constexpr int S = TPI * L;
    static constexpr uint32_t GMASK = (TPI == 32) ? 0xffffffffu : ((1u << TPI) - 1u);

    uint32_t n[L];   // this lane's limbs of the modulus
    uint32_t np0;    // -n^-1 mod 2^32
    int t;           // lane index inside the group
    int gshift;      // bit position of the group's lane 0 inside the warp

    __device__ __forceinline__ void init(const uint32_t* __restrict__ nmod, uint32_t np0_) {
        const int lane = threadIdx.x & 31;
        t = lane & (TPI - 1);
        gshift = lane & ~(TPI - 1);
        np0 = np0_;
#pragma unroll
        for (int k = 0; k < L; ++k) n[k] = nmod[t * L + k];
    }

    __device__ __forceinline__ uint32_t gballot(bool pred) const {
        return (__ballot_sync(FULL_MASK, pred) >> gshift) & GMASK;
    }

    // r (L limbs + overflow word ov at position L of this lane) -> canonical
    // residue in [0, n): resolves the cross-lane carries, then subtracts n if
    // the value is >= n.  Precondition: total value < 2n.
    __device__ __forceinline__ void resolve_reduce(uint32_t (&r)[L], uint32_t ov) {
        // 1. hand the overflow word to the next lane and ripple it in.
        uint32_t in = __shfl_up_sync(FULL_MASK, ov, 1, TPI);
        if (t == 0) in = 0;
        uint32_t cy;
        add_cc(r[0], r[0], in);
#pragma unroll
        for (int k = 1; k < L; ++k) addc_cc(r[k], r[k], 0);
        addc(cy, 0, 0);
        uint32_t all1 = r[0];
#pragma unroll
        for (int k = 1; k < L; ++k) all1 &= r[k];
        uint32_t g = gballot(cy != 0);
        uint32_t p = gballot(all1 == 0xffffffffu);
        uint64_t ci = lookahead(g, p);
        uint32_t cin = (uint32_t)(ci >> t) & 1u;
        add_cc(r[0], r[0], cin);
#pragma unroll
        for (int k = 1; k < L; ++k) addc_cc(r[k], r[k], 0);
        // carry out of the whole number: top lane's own overflow word + look-ahead carry
        uint32_t top_ov = __shfl_sync(FULL_MASK, ov, TPI - 1, TPI);
        uint32_t overflow = top_ov + ((uint32_t)(ci >> TPI) & 1u);

        // 2. d = r - n with cross-lane borrow resolution.
        uint32_t d[L], bw;
        sub_cc(d[0], r[0], n[0]);
#pragma unroll
        for (int k = 1; k < L; ++k) subc_cc(d[k], r[k], n[k]);
        subc(bw, 0, 0);          // 0 - 0 - borrow -> 0xffffffff if borrow
        uint32_t any = d[0];
#pragma unroll
        for (int k = 1; k < L; ++k) any |= d[k];
        uint32_t bg = gballot(bw != 0);
        uint32_t bp = gballot(any == 0);
        uint64_t bi = lookahead(bg, bp);
        uint32_t bin = (uint32_t)(bi >> t) & 1u;
        sub_cc(d[0], d[0], bin);
#pragma unroll
        for (int k = 1; k < L; ++k) subc_cc(d[k], d[k], 0);
        uint32_t final_borrow = (uint32_t)(bi >> TPI) & 1u;
        // r >= n  <=>  overflow word set or no final borrow
        bool take = (overflow != 0) || (final_borrow == 0);
#pragma unroll
        for (int k = 0; k < L; ++k) r[k] = take ? d[k] : r[k];
    }

    // r = a * b * R^-1 mod n,  R = 2^(32*S).  Requires a < R, b < n (or a*b < n*R).
    __device__ __forceinline__ void mul(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
        uint32_t E[L], O[L], Ec = 0, Oc = 0;
#pragma unroll
        for (int k = 0; k < L; ++k) { E[k] = 0; O[k] = 0; }

#pragma unroll 1
        for (int u = 0; u < TPI; ++u) {
#pragma unroll
            for (int k = 0; k < L; ++k) {
                const uint32_t bj = __shfl_sync(FULL_MASK, b[k], u, TPI);
                // --- one-limb shift of the previous state, fused into a*bj ---
                uint32_t recv = __shfl_down_sync(FULL_MASK, E[0], 1, TPI);
                if (t == TPI - 1) recv = 0;
                const uint64_t top = (uint64_t)Ec + recv;
                const uint32_t top_lo = (uint32_t)top, top_hi = (uint32_t)(top >> 32);
                uint32_t nE[L], nO[L], nEc, nOc;
                // new even array = old odd array; the old E[1] drops onto its limb 0
                add_cc(nE[0], O[0], E[1]);
                // new odd array = old even array >> 64, plus a[odd]*bj
#pragma unroll
                for (int j = 0; j < L - 2; j += 2) {
                    madc_lo_cc(nO[j], a[j + 1], bj, E[j + 2]);
                    madc_hi_cc(nO[j + 1], a[j + 1], bj, E[j + 3]);
                }
                madc_lo_cc(nO[L - 2], a[L - 1], bj, top_lo);
                madc_hi_cc(nO[L - 1], a[L - 1], bj, top_hi);
                addc(nOc, 0, 0);
                // a[even]*bj onto the new even array
                mad_lo_cc(nE[0], a[0], bj, nE[0]);
                madc_hi_cc(nE[1], a[0], bj, O[1]);
#pragma unroll
                for (int j = 2; j < L; j += 2) {
                    madc_lo_cc(nE[j], a[j], bj, O[j]);
                    madc_hi_cc(nE[j + 1], a[j], bj, O[j + 1]);
                }
                addc(nEc, Oc, 0);
                // --- quotient digit from the group's lowest limb ---
                const uint32_t q = __shfl_sync(FULL_MASK, nE[0] * np0, 0, TPI);
                // --- n*q ---
                mad_lo_cc(nE[0], n[0], q, nE[0]);
                madc_hi_cc(nE[1], n[0], q, nE[1]);
#pragma unroll
                for (int j = 2; j < L; j += 2) {
                    madc_lo_cc(nE[j], n[j], q, nE[j]);
                    madc_hi_cc(nE[j + 1], n[j], q, nE[j + 1]);
                }
                addc(nEc, nEc, 0);
                mad_lo_cc(nO[0], n[1], q, nO[0]);
                madc_hi_cc(nO[1], n[1], q, nO[1]);
#pragma unroll
                for (int j = 2; j < L; j += 2) {
                    madc_lo_cc(nO[j], n[j + 1], q, nO[j]);
                    madc_hi_cc(nO[j + 1], n[j + 1], q, nO[j + 1]);
                }
                addc(nOc, nOc, 0);
#pragma unroll
                for (int j = 0; j < L; ++j) { E[j] = nE[j]; O[j] = nO[j]; }
                Ec = nEc; Oc = nOc;
            }
        }
        // final one-limb shift and merge of the two arrays
        uint32_t recv = __shfl_down_sync(FULL_MASK, E[0], 1, TPI);
        if (t == TPI - 1) recv = 0;
        const uint64_t top = (uint64_t)Ec + recv;
        const uint32_t top_lo = (uint32_t)top, top_hi = (uint32_t)(top >> 32);
        uint32_t ov;
        add_cc(r[0], O[0], E[1]);
#pragma unroll
        for (int k = 1; k < L - 1; ++k) addc_cc(r[k], O[k], E[k + 1]);
        addc_cc(r[L - 1], O[L - 1], top_lo);
        addc(ov, Oc, top_hi);
        resolve_reduce(r, ov);
    }

    // r = (a + b) mod n, a, b < n
    __device__ __forceinline__ void add(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
        uint32_t ov;
        add_cc(r[0], a[0], b[0]);
#pragma unroll
        for (int k = 1; k < L; ++k) addc_cc(r[k], a[k], b[k]);
        addc(ov, 0, 0);
        resolve_reduce(r, ov);
    }

    // r = (a - b) mod n, a, b < n
    __device__ __forceinline__ void sub(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
        uint32_t d[L], bw;
        sub_cc(d[0], a[0], b[0]);
#pragma unroll
        for (int k = 1; k < L; ++k) subc_cc(d[k], a[k], b[k]);
        subc(bw, 0, 0);
        uint32_t any = d[0];
#pragma unroll
        for (int k = 1; k < L; ++k) any |= d[k];
        const uint32_t bg = gballot(bw != 0);
        const uint32_t bp = gballot(any == 0);
        const uint64_t bi = lookahead(bg, bp);
        const uint32_t bin = (uint32_t)(bi >> t) & 1u;
        sub_cc(d[0], d[0], bin);
#pragma unroll
        for (int k = 1; k < L; ++k) subc_cc(d[k], d[k], 0);
        const bool negative = ((uint32_t)(bi >> TPI) & 1u) != 0;
        // a < b: add n back (the value is d + n - 2^(32S), exact in S limbs)
        uint32_t s[L], cy;
        add_cc(s[0], d[0], n[0]);
#pragma unroll
        for (int k = 1; k < L; ++k) addc_cc(s[k], d[k], n[k]);
        addc(cy, 0, 0);
        uint32_t all1 = s[0];
#pragma unroll
        for (int k = 1; k < L; ++k) all1 &= s[k];
        const uint32_t g = gballot(cy != 0);
        const uint32_t p = gballot(all1 == 0xffffffffu);
        const uint64_t ci = lookahead(g, p);
        const uint32_t cin = (uint32_t)(ci >> t) & 1u;
        add_cc(s[0], s[0], cin);
#pragma unroll
        for (int k = 1; k < L; ++k) addc_cc(s[k], s[k], 0);
#pragma unroll
        for (int k = 0; k < L; ++k) r[k] = negative ? s[k] : d[k];
    }
};

}  // namespace pgpu
