// Warp-cooperative Montgomery arithmetic on the FP64 pipe of sm_100a: 52-bit limbs held as doubles.
//
// Why: B200 issues 64 DFMA per SM per clock but only 32 IMAD.WIDE.U32 (tools/dfma_peak.cu, tools/imad_peak.cu:
// 18.46 T DFMA/s against 9.25 T IMAD.WIDE/s).  A 52x52-bit limb product costs two DFMA.RZ and one DADD
//     hi = fma_rz(a, b, 2^104)            -> 2^104 + floor(a*b / 2^52) * 2^52      (exact: ulp(2^104) = 2^52)
//     lo = fma_rz(a, b, 2^104 + 2^52 - hi) -> 2^52 + (a*b mod 2^52)                 (exact: in [2^52, 2^53))
// and the two results are accumulated as 64-bit INTEGERS on their bit patterns (IADD3 on the ALU pipe): the exponent
// fields are constants, the mantissa fields add up like integers.  2704 bit^2 per 3 FP64-pipe slots against 1024 bit^2
// per IMAD.WIDE slot on paper; measured, the carry adds do not hide behind the FP64 pipe and the two routes tie at
// 9.1 vs 9.47 Pbit^2/s (profiles/r02_dfma_peak.json, r02_fp64_experiments.md): this kernel wins 3-17 % by shape, not 1.6x.
//
// Layout: a residue of S52 = TPI*L limbs of 52 bits, lane t of the group keeps limbs [t*L, t*L+L) as doubles (exact
// integers < 2^52).  R = 2^(52*S52) > 4n, so values are kept lazily in [0, 2n) and no multiplication ends with a
// conditional subtraction; records are made canonical when they are stored.  CIOS as in mont.cuh: per limb b_j of the
// second operand the group accumulates a*b_j, derives the quotient digit q from column 0, accumulates n*q and shifts one
// column down (column 0 of lane t+1 moves into the top column of lane t).
//
// Exponent-field bookkeeping.  Every hi pattern carries 0x467<<52, every lo pattern 0x433<<52.  Both are multiples of
// 2^52, so the low 52 bits of a column -- all the quotient digit needs -- are never disturbed.  A column lives in a lane
// for exactly L rows (it enters as the spill column above the lane's top product and leaves through column 0) and
// collects L*(2*0x433 + 2*0x467)<<52 on the way; it is created holding minus that constant, so it leaves the lane
// clean: the carry out of column 0 and the column handed to the neighbouring lane are plain integers.  The L columns
// left over at the end get a per-position correction (high word only).
//
// This is the multiplier behind the reference's mpz_powm calls (e.g. /root/reference/paillier.go:213-216,
// thresholdkey.go:199) for moduli of 2048 bits and more.
#pragma once
#include <cstdint>
#include "mont.cuh"

namespace pgpu {

__device__ __forceinline__ double u52_to_double(uint64_t v) {
    return __longlong_as_double((long long)(v | 0x4330000000000000ull)) - 0x1p52;
}
__device__ __forceinline__ uint64_t double_to_u52(double d) {
    return (uint64_t)__double_as_longlong(d + 0x1p52) & ((1ull << 52) - 1);
}

// bit patterns of the high and low halves of the 104-bit product a*b (a, b integers < 2^52 held as doubles)
__device__ __forceinline__ void split52(double a, double b, uint64_t& h, uint64_t& l) {
    const double hd = __fma_rz(a, b, 0x1p104);
    const double ld = __fma_rz(a, b, __dsub_rn(0x1.0000000000001p104, hd));
    h = (uint64_t)__double_as_longlong(hd);
    l = (uint64_t)__double_as_longlong(ld);
}

#ifndef PGPU52_ADD
#define PGPU52_ADD 0      // 0: three-input 64-bit accumulations (IADD3 with two carries), 1: two-input adds (tools/mont52_test A/B)
#endif
// acc += x as one add.cc / addc pair the compiler cannot fold into a three-input add
__device__ __forceinline__ void add64_2in(uint64_t& acc, uint64_t x) {
    uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
    asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"((uint32_t)x), "r"((uint32_t)(x >> 32)));
    acc = ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void acc2(uint64_t& acc, uint64_t x, uint64_t y) {
#if PGPU52_ADD == 1
    add64_2in(acc, x); add64_2in(acc, y);
#else
    acc += x + y;
#endif
}

template <int TPI, int L, int S32_>
struct Mont52 {
    static_assert(TPI >= 2 && TPI <= 32 && (TPI & (TPI - 1)) == 0, "TPI must be a power of two");
    static_assert(L >= 2, "at least two limbs per lane");
    static constexpr int S52 = TPI * L;
    static constexpr int S32 = S32_;                       // 32-bit limbs of a record of this modulus
    static_assert(52 * S52 >= 32 * S32 + 2, "R = 2^(52*S52) must exceed 4n");
    static constexpr int RBITS = 52 * S52;
    static constexpr uint32_t GMASK = (TPI == 32) ? 0xffffffffu : ((1u << TPI) - 1u);
    static constexpr uint64_t M52 = (1ull << 52) - 1;
    static constexpr int TBL = 2 * L;                      // 32-bit words of one table entry per lane
    using elem = double;

    // exponent fields (high 32-bit words of the patterns)
    static constexpr uint32_t BH = 0x46700000u, BL = 0x43300000u;
    static constexpr uint32_t ROWB = 2u * BL + 2u * BH;    // what a column at local position >= 1 collects per row

    double n[L];     // this lane's limbs of the modulus
    uint64_t np;     // -n^-1 mod 2^52
    int t;           // lane index inside the group
    int gshift;      // bit position of the group's lane 0 inside the warp

    __device__ __forceinline__ uint32_t gballot(bool pred) const {
        return (__ballot_sync(FULL_MASK, pred) >> gshift) & GMASK;
    }

    __device__ __forceinline__ void init(const uint32_t* __restrict__ nmod, uint32_t np0_) {
        const int lane = threadIdx.x & 31;
        t = lane & (TPI - 1);
        gshift = lane & ~(TPI - 1);
        const uint64_t n64 = (uint64_t)nmod[0] | ((uint64_t)nmod[1] << 32);
        uint64_t inv = (uint64_t)(0u - np0_);          // n^-1 mod 2^32
        inv *= 2ull - n64 * inv;                       // mod 2^64
        np = (0ull - inv) & M52;
        load_rec(n, nmod, S32);
    }

    // ---------------------------------------------------------------- records of 32-bit limbs <-> 52-bit limbs
    // Lane t owns bits [52*L*t, 52*L*(t+1)) of the integer: the words covering them are fetched, shifted to bit 0 of a
    // "lane string" (dynamic funnel shift), and the limbs are cut out at static positions.
    static constexpr int NQ = (52 * (L - 1)) / 32 + 3;     // aligned 32-bit pieces the limb extraction touches
    __device__ __forceinline__ void load_rec(double (&x)[L], const uint32_t* __restrict__ p, uint32_t lim) const {
        const uint32_t bit0 = 52u * L * (uint32_t)t;
        const uint32_t w0 = bit0 >> 5, o = bit0 & 31u;
        uint32_t P[NQ + 1], Q[NQ];
#pragma unroll
        for (int r = 0; r <= NQ; ++r) { const uint32_t idx = w0 + r; P[r] = idx < lim ? __ldg(p + idx) : 0u; }
#pragma unroll
        for (int r = 0; r < NQ; ++r) Q[r] = __funnelshift_r(P[r], P[r + 1], o);
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const int s = 52 * k, wi = s >> 5, sh = s & 31;
            const uint32_t lo = __funnelshift_r(Q[wi], Q[wi + 1], sh);
            const uint32_t hi = __funnelshift_r(Q[wi + 1], Q[wi + 2], sh) & 0xfffffu;
            x[k] = u52_to_double((uint64_t)lo | ((uint64_t)hi << 32));
        }
    }

    // canonical integer limbs X (< 2^52 each, value < 2^(32*lim)) -> record
    static constexpr int NS = (52 * L + 31) / 32 + 2;      // pieces of the lane string incl. the neighbour's limb 0
    __device__ __forceinline__ void store_limbs(uint32_t* __restrict__ p, const uint64_t (&X)[L], uint32_t lim) const {
        uint64_t nx = __shfl_down_sync(FULL_MASK, X[0], 1, TPI);
        if (t == TPI - 1) nx = 0;
        uint32_t Sg[NS + 1];
#pragma unroll
        for (int r = 0; r <= NS; ++r) {
            // bits [32r, 32r+32) of sum X[k] 2^(52k) + nx 2^(52L)
            const int b = 32 * r;
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k <= L; ++k) {
                const int lo = 52 * k, hi = lo + 52;           // limb k covers [lo, hi)
                if (hi > b && lo < b + 32) {
                    const uint64_t v = (k < L) ? X[k] : nx;
                    if (lo >= b) w |= (uint32_t)(v << (lo - b));
                    else w |= (uint32_t)(v >> (b - lo));
                }
            }
            Sg[r] = w;
        }
        const uint32_t bit0 = 52u * L * (uint32_t)t, bit1 = bit0 + 52u * L;
        const uint32_t jb = (bit0 + 31u) >> 5, je = (t == TPI - 1) ? 0xffffffffu : ((bit1 + 31u) >> 5);
        const uint32_t o2 = 32u * jb - bit0;
#pragma unroll
        for (int r = 0; r < NS; ++r) {
            const uint32_t j = jb + r;
            if (j < je && j < lim) p[j] = __funnelshift_r(Sg[r], Sg[r + 1], o2);
        }
    }

    // ---------------------------------------------------------------- integer-limb helpers (limbs < 2^52 unless noted)
    // columns (< 2^63) -> normalised limbs; the carry out of the whole group is dropped (callers guarantee value < R)
    __device__ __forceinline__ void normalize(uint64_t (&C)[L]) const {
        uint64_t c = 0;
#pragma unroll
        for (int k = 0; k < L; ++k) { const uint64_t v = C[k] + c; C[k] = v & M52; c = v >> 52; }
        uint32_t cin = __shfl_up_sync(FULL_MASK, (uint32_t)c, 1, TPI);     // < 2^12
        if (t == 0) cin = 0;
        uint64_t c2 = cin;
#pragma unroll
        for (int k = 0; k < L; ++k) { const uint64_t v = C[k] + c2; C[k] = v & M52; c2 = v >> 52; }
        uint64_t all1 = C[0];
#pragma unroll
        for (int k = 1; k < L; ++k) all1 &= C[k];
        const uint32_t g = gballot(c2 != 0), p = gballot(all1 == M52);
        const uint64_t ci = lookahead(g, p);
        uint64_t c3 = (ci >> t) & 1u;
#pragma unroll
        for (int k = 0; k < L; ++k) { const uint64_t v = C[k] + c3; C[k] = v & M52; c3 = v >> 52; }
    }

    // d = x - y mod 2^(52*S52); returns true when x < y
    __device__ __forceinline__ bool sub_limbs(uint64_t (&d)[L], const uint64_t (&x)[L], const uint64_t (&y)[L]) const {
        uint64_t bw = 0;
#pragma unroll
        for (int k = 0; k < L; ++k) { const uint64_t v = x[k] - y[k] - bw; d[k] = v & M52; bw = v >> 63; }
        uint64_t any = d[0];
#pragma unroll
        for (int k = 1; k < L; ++k) any |= d[k];
        const uint32_t g = gballot(bw != 0), p = gballot(any == 0);
        const uint64_t bi = lookahead(g, p);
        uint64_t b2 = (bi >> t) & 1u;
#pragma unroll
        for (int k = 0; k < L; ++k) { const uint64_t v = d[k] - b2; d[k] = v & M52; b2 = v >> 63; }
        return ((bi >> TPI) & 1u) != 0;
    }

    __device__ __forceinline__ void to_limbs(uint64_t (&X)[L], const double (&x)[L]) const {
#pragma unroll
        for (int k = 0; k < L; ++k) X[k] = double_to_u52(x[k]);
    }
    __device__ __forceinline__ void from_limbs(double (&x)[L], const uint64_t (&X)[L]) const {
#pragma unroll
        for (int k = 0; k < L; ++k) x[k] = u52_to_double(X[k]);
    }
    // k * n as normalised limbs (k = 1, 2)
    __device__ __forceinline__ void n_times(uint64_t (&X)[L], int k) const {
#pragma unroll
        for (int i = 0; i < L; ++i) X[i] = double_to_u52(n[i]) * (uint64_t)k;
        normalize(X);
    }
    // X < 2m -> X mod m for m = k*n
    __device__ __forceinline__ void cond_sub(uint64_t (&X)[L], int k) const {
        uint64_t m[L], d[L];
        n_times(m, k);
        const bool less = sub_limbs(d, X, m);
#pragma unroll
        for (int i = 0; i < L; ++i) X[i] = less ? X[i] : d[i];
    }

    // value in [0, 2n) -> canonical record
    __device__ __forceinline__ void store_rec(uint32_t* __restrict__ p, const double (&x)[L], uint32_t lim) const {
        uint64_t X[L];
        to_limbs(X, x);
        cond_sub(X, 1);
        store_limbs(p, X, lim);
    }

    // table entries keep the lazy double form
    __device__ __forceinline__ void load_tbl(double (&x)[L], const uint32_t* __restrict__ p) const {
        if constexpr (L % 2 == 0) {
            const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
            for (int k = 0; k < L / 2; ++k) { const double2 v = q[k]; x[2 * k] = v.x; x[2 * k + 1] = v.y; }
        } else {
            const double* q = reinterpret_cast<const double*>(p);
#pragma unroll
            for (int k = 0; k < L; ++k) x[k] = q[k];
        }
    }
    __device__ __forceinline__ void store_tbl(uint32_t* __restrict__ p, const double (&x)[L]) const {
        if constexpr (L % 2 == 0) {
            double2* q = reinterpret_cast<double2*>(p);
#pragma unroll
            for (int k = 0; k < L / 2; ++k) q[k] = make_double2(x[2 * k], x[2 * k + 1]);
        } else {
            double* q = reinterpret_cast<double*>(p);
#pragma unroll
            for (int k = 0; k < L; ++k) q[k] = x[k];
        }
    }

    // ---------------------------------------------------------------- r = a * b * R^-1 mod n, lazily in [0, 2n)
    __device__ __forceinline__ void mul(double (&r)[L], const double (&a)[L], const double (&b)[L]) {
#if PGPU52_ADD == 2
        mul_cf(r, a, b);
        return;
#endif
        uint64_t C[L + 1];
#pragma unroll
        for (int k = 0; k < L; ++k) C[k] = (uint64_t)(0u - ((uint32_t)k * ROWB + 2u * BL)) << 32;
        constexpr uint64_t KSP = (uint64_t)(0u - (uint32_t)L * ROWB) << 32;
        C[L] = 0;

#pragma unroll 1
        for (int u = 0; u < TPI; ++u) {
#pragma unroll
            for (int k = 0; k < L; ++k) {
                const double bj = __shfl_sync(FULL_MASK, b[k], u, TPI);
                uint64_t h[L], l[L];
                // a * b_j
#pragma unroll
                for (int i = 0; i < L; ++i) split52(a[i], bj, h[i], l[i]);
                C[0] += l[0];
#pragma unroll
                for (int i = 1; i < L; ++i) acc2(C[i], l[i], h[i - 1]);
                C[L] = KSP + h[L - 1];
                // quotient digit from the group's column 0
                const uint64_t q = __shfl_sync(FULL_MASK, ((C[0] & M52) * np) & M52, 0, TPI);
                const double qd = u52_to_double(q);
                // n * q
#pragma unroll
                for (int i = 0; i < L; ++i) split52(n[i], qd, h[i], l[i]);
                C[0] += l[0];
#pragma unroll
                for (int i = 1; i < L; ++i) acc2(C[i], l[i], h[i - 1]);
                C[L] += h[L - 1];
                // one column down: column 0 is complete (and clean of exponent fields)
                uint64_t recv = __shfl_down_sync(FULL_MASK, C[0], 1, TPI);
                if (t == TPI - 1) recv = 0;
                const uint64_t carry = (t == 0) ? (C[0] >> 52) : 0ull;
#pragma unroll
                for (int i = 0; i < L - 1; ++i) C[i] = C[i + 1];
                C[L - 1] = C[L] + recv;
                C[0] += carry;
            }
        }
        uint64_t X[L];
#pragma unroll
        for (int k = 0; k < L; ++k) X[k] = C[k] + ((uint64_t)((uint32_t)(k + 1) * ROWB - 2u * BH) << 32);
        normalize(X);
        from_limbs(r, X);
    }

    // ---------------------------------------------------------------- the same product with carry-free accumulators
    // Carry-propagating adds (IADD3 with carry-out predicates, IADD3.X) do not issue in the shadow of the FP64 pipe the way
    // plain adds do (tools/pipe_probe.cu).  Here every 64-bit column is three 32-bit words updated by plain adds only:
    //   lo = sum of the low words mod 2^32,  hi = sum of the high words mod 2^32,  ap = sum of (low word >> 16).
    // The carries the low words lost are recovered when a column is read: the true low sum T is congruent to lo mod 2^32
    // and lies in [ap*2^16, ap*2^16 + N*2^16) for N terms (N*2^16 < 2^32), so T = lo + 2^32*k with
    // k = (ap*2^16 + (2^32 - 1 - lo)) >> 32.
    struct Col3 { uint32_t lo, hi, ap; };
    static __device__ __forceinline__ uint64_t col_value(const Col3& c) {
        const uint64_t k = ((((uint64_t)c.ap) << 16) + (uint64_t)(~c.lo)) >> 32;
        return (uint64_t)c.lo | ((uint64_t)(c.hi + (uint32_t)k) << 32);
    }
    static __device__ __forceinline__ void col_add1(Col3& c, uint64_t x) {
        const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
        c.lo += xl; c.hi += xh; c.ap += xl >> 16;
    }
    static __device__ __forceinline__ void col_add2(Col3& c, uint64_t x, uint64_t y) {
        const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32), yl = (uint32_t)y, yh = (uint32_t)(y >> 32);
        c.lo = c.lo + xl + yl; c.hi = c.hi + xh + yh; c.ap = c.ap + (xl >> 16) + (yl >> 16);
    }

    __device__ __forceinline__ void mul_cf(double (&r)[L], const double (&a)[L], const double (&b)[L]) {
        Col3 C[L + 1];
#pragma unroll
        for (int k = 0; k < L; ++k) { C[k].lo = 0; C[k].ap = 0; C[k].hi = 0u - ((uint32_t)k * ROWB + 2u * BL); }
        constexpr uint32_t KSPH = 0u - (uint32_t)L * ROWB;
        C[L].lo = C[L].hi = C[L].ap = 0;

#pragma unroll 1
        for (int u = 0; u < TPI; ++u) {
#pragma unroll
            for (int k = 0; k < L; ++k) {
                const double bj = __shfl_sync(FULL_MASK, b[k], u, TPI);
                uint64_t h[L], l[L];
#pragma unroll
                for (int i = 0; i < L; ++i) split52(a[i], bj, h[i], l[i]);
                col_add1(C[0], l[0]);
#pragma unroll
                for (int i = 1; i < L; ++i) col_add2(C[i], l[i], h[i - 1]);
                C[L].lo = (uint32_t)h[L - 1]; C[L].hi = KSPH + (uint32_t)(h[L - 1] >> 32); C[L].ap = (uint32_t)h[L - 1] >> 16;
                const uint64_t v0 = col_value(C[0]);
                const uint64_t q = __shfl_sync(FULL_MASK, ((v0 & M52) * np) & M52, 0, TPI);
                const double qd = u52_to_double(q);
#pragma unroll
                for (int i = 0; i < L; ++i) split52(n[i], qd, h[i], l[i]);
                col_add1(C[0], l[0]);
#pragma unroll
                for (int i = 1; i < L; ++i) col_add2(C[i], l[i], h[i - 1]);
                col_add1(C[L], h[L - 1]);
                const uint64_t done = col_value(C[0]);
                uint64_t recv = __shfl_down_sync(FULL_MASK, done, 1, TPI);
                if (t == TPI - 1) recv = 0;
                const uint32_t carry = (t == 0) ? (uint32_t)(done >> 52) : 0u;
#pragma unroll
                for (int i = 0; i < L - 1; ++i) C[i] = C[i + 1];
                C[L - 1] = C[L];
                col_add1(C[L - 1], recv);
                C[0].lo += carry;            // < 2^12: nothing for ap
            }
        }
        uint64_t X[L];
#pragma unroll
        for (int k = 0; k < L; ++k) X[k] = col_value(C[k]) + ((uint64_t)((uint32_t)(k + 1) * ROWB - 2u * BH) << 32);
        normalize(X);
        from_limbs(r, X);
    }

    __device__ __forceinline__ void sqr(double (&r)[L], const double (&a)[L]) { mul(r, a, a); }

    // r = a + b mod n (lazy): a, b < 2n -> r < 2n
    __device__ __forceinline__ void add(double (&r)[L], const double (&a)[L], const double (&b)[L]) {
        uint64_t X[L];
#pragma unroll
        for (int k = 0; k < L; ++k) X[k] = double_to_u52(a[k]) + double_to_u52(b[k]);
        normalize(X);
        cond_sub(X, 2);
        from_limbs(r, X);
    }

    // r = a - b mod n (lazy): a + 2n - b in (0, 4n) -> r < 2n
    __device__ __forceinline__ void sub(double (&r)[L], const double (&a)[L], const double (&b)[L]) {
        uint64_t X[L], Y[L], D[L];
        n_times(X, 2);
#pragma unroll
        for (int k = 0; k < L; ++k) X[k] += double_to_u52(a[k]);
        normalize(X);
        to_limbs(Y, b);
        sub_limbs(D, X, Y);
        cond_sub(D, 2);
        from_limbs(r, D);
    }
};

}  // namespace pgpu
