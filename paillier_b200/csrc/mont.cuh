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

#ifndef PGPU_FENCE
#define PGPU_FENCE 3   // rows of a block product in flight (see Mont::prod_row)
#endif

namespace pgpu {

constexpr unsigned FULL_MASK = 0xffffffffu;

#ifndef PGPU_HOST_EMULATION
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
#else
// ---- the same primitives for tests/cpp/mont_host_test.cpp, which runs THIS header on CPU threads (one per lane of an
// emulated warp, tests/cpp/cuda_host_shim.h): the CC.CF flag of a thread is a thread-local variable.
inline thread_local uint32_t pgpu_cc = 0;
inline void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) {
    const uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c; d = (uint32_t)s; pgpu_cc = (uint32_t)(s >> 32);
}
inline void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) {
    const uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c + pgpu_cc; d = (uint32_t)s; pgpu_cc = (uint32_t)(s >> 32);
}
inline void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) {
    const uint64_t s = (((uint64_t)a * b) >> 32) + c + pgpu_cc; d = (uint32_t)s; pgpu_cc = (uint32_t)(s >> 32);
}
inline void add_cc(uint32_t& d, uint32_t a, uint32_t b) { const uint64_t s = (uint64_t)a + b; d = (uint32_t)s; pgpu_cc = (uint32_t)(s >> 32); }
inline void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { const uint64_t s = (uint64_t)a + b + pgpu_cc; d = (uint32_t)s; pgpu_cc = (uint32_t)(s >> 32); }
inline void addc(uint32_t& d, uint32_t a, uint32_t b) { d = a + b + pgpu_cc; }
inline void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { const uint32_t bw = a < b; d = a - b; pgpu_cc = bw; }
inline void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { const uint32_t bw = (uint64_t)a < (uint64_t)b + pgpu_cc; d = a - b - pgpu_cc; pgpu_cc = bw; }
inline void subc(uint32_t& d, uint32_t a, uint32_t b) { d = a - b - pgpu_cc; }
#endif

// Carry look-ahead over a group: g = lanes that generate a carry, p = lanes
// that propagate one (mutually exclusive).  Returns the mask of lanes that
// receive a carry-in; bit TPI is the carry out of the whole group.
__device__ __forceinline__ uint64_t lookahead(uint32_t g, uint32_t p) {
    return ((uint64_t)(g | p) + g) ^ p;
}

// Shapes that use the dedicated squaring.  Measured on B200 (profiles/r01_sqr_experiments.md): with L <= 16 the kernel
// runs 4 warps per scheduler and the squaring's 19 % fewer multiplies turn into +6 % decrypt throughput; with L = 32
// only 2 warps per scheduler fit, every warp instruction then costs ~6.5 cycles of issue latency, and a loop needs
// >= 81 % IMAD.WIDE in its instruction mix to saturate the multiplier pipe -- mul() has 84 %, the squaring's reduction
// loop (half the multiplies per row, same per-row bookkeeping) 73 % -- so mul(a, a) stays faster there.
template <int TPI_, int L_> struct SqrShape { static constexpr bool value = (TPI_ == 4) && (L_ % 8 == 0) && (L_ <= 16); };

#ifndef PGPU_HOST_EMULATION
__device__ __forceinline__ uint4 lds_v4_volatile(const uint4* p) {
    uint4 v;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
#else
inline uint4 lds_v4_volatile(const uint4* p) { return *p; }
#endif

// NSM: the modulus limbs are kept in shared memory and fetched at every use instead of occupying L registers
// for the whole kernel (needed by the dedicated squaring, whose product phase has no use for them).
template <int TPI, int L, bool NSM = false>
struct Mont {
    static_assert((L % 2) == 0, "L must be even");
    static_assert(TPI >= 1 && TPI <= 32 && (TPI & (TPI - 1)) == 0, "TPI must be a power of two");
    static constexpr int S = TPI * L;
    static constexpr uint32_t GMASK = (TPI == 32) ? 0xffffffffu : ((1u << TPI) - 1u);

    uint32_t n[L];   // this lane's limbs of the modulus
    uint32_t np0;    // -n^-1 mod 2^32
    uint64_t np64;   // -n^-1 mod 2^64 (two quotient digits at a time in the squaring's reduction)
    int t;           // lane index inside the group
    int gshift;      // bit position of the group's lane 0 inside the warp

    __device__ __forceinline__ void init(const uint32_t* __restrict__ nmod, uint32_t np0_) {
        const int lane = threadIdx.x & 31;
        t = lane & (TPI - 1);
        gshift = lane & ~(TPI - 1);
        np0 = np0_;
        {
            const uint64_t n64 = (uint64_t)nmod[0] | ((uint64_t)nmod[1] << 32);
            uint64_t inv = (uint64_t)(0u - np0_);          // n^-1 mod 2^32
            inv *= 2ull - n64 * inv;                       // mod 2^64
            np64 = 0ull - inv;
        }
#pragma unroll
        for (int k = 0; k < L; ++k) n[k] = nmod[t * L + k];
    }

    // this lane's limbs of the modulus
    __device__ __forceinline__ void fetch_n(uint32_t (&nn)[L]) const {
        if constexpr (NSM) {
#pragma unroll
            for (int k4 = 0; k4 < L / 4; ++k4) {
                const uint4 q = lds_v4_volatile(sn + k4 * 32 + wl);
                nn[4 * k4] = q.x; nn[4 * k4 + 1] = q.y; nn[4 * k4 + 2] = q.z; nn[4 * k4 + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < L; ++k) nn[k] = n[k];
        }
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
        uint32_t d[L], bw, nn[L];
        fetch_n(nn);
        sub_cc(d[0], r[0], nn[0]);
#pragma unroll
        for (int k = 1; k < L; ++k) subc_cc(d[k], r[k], nn[k]);
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
        uint32_t E[L], O[L], Ec = 0, Oc = 0, nn[L];
        fetch_n(nn);
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
                mad_lo_cc(nE[0], nn[0], q, nE[0]);
                madc_hi_cc(nE[1], nn[0], q, nE[1]);
#pragma unroll
                for (int j = 2; j < L; j += 2) {
                    madc_lo_cc(nE[j], nn[j], q, nE[j]);
                    madc_hi_cc(nE[j + 1], nn[j], q, nE[j + 1]);
                }
                addc(nEc, nEc, 0);
                mad_lo_cc(nO[0], nn[1], q, nO[0]);
                madc_hi_cc(nO[1], nn[1], q, nO[1]);
#pragma unroll
                for (int j = 2; j < L; j += 2) {
                    madc_lo_cc(nO[j], nn[j + 1], q, nO[j]);
                    madc_hi_cc(nO[j + 1], nn[j + 1], q, nO[j + 1]);
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


    // ------------------------------------------------------------------------------------------------
    // Dedicated Montgomery squaring (TPI == 4): r = a * a * R^-1 mod n with ~19 % fewer multiply-accumulates
    // than mul(a, a).
    //
    // Phase 1 (product): a = A0 + A1 X + A2 X^2 + A3 X^3, X = 2^(32 L), lane t holding A_t.  The ten block
    // products of a^2 are shared out evenly: lane t computes O_t = A_t * A_(t+1 mod 4) (counted twice in a^2),
    // one half of the remaining pair product (lanes 0,2 split A_0 * A_2; lanes 1,3 split A_1 * A_3; counted
    // twice) and D_t = A_t^2.  Every product is a 2L-limb buffer whose low half belongs to slot X^k and whose
    // high half belongs to slot X^(k+1); lane r owns slots X^r (low half of a^2) and X^(r+4) (high half), which
    // is exactly the layout the reduction wants.  The buffers are exchanged through shared memory
    // ([uint4 row][lane of the warp] so that all accesses are conflict free).
    // Phase 2 (reduction): the low half goes through the same sliding even/odd window as mul(), with only the
    // n*q chains, and the high half is added at the end.
    // ------------------------------------------------------------------------------------------------
    static constexpr bool HAS_SQR = NSM && SqrShape<TPI, L>::value;
    static constexpr int SQR_A4 = L / 4;            // uint4 rows holding one operand per lane
    static constexpr int SQR_X4 = 2 * L / 4;        // uint4 rows of one exchange buffer per lane
    static constexpr int SQR_ROWS = SQR_A4 + SQR_X4 + 1 + SQR_A4 + SQR_X4;   // + one row of zeros + the modulus + accumulator stash
    static constexpr size_t SQR_SMEM_PER_WARP = (size_t)SQR_ROWS * 32 * 16;

    uint4* sa;      // [SQR_A4][32] operand rows of this warp
    uint4* sx;      // [SQR_X4][32] exchange rows
    uint4* sz;      // [1][32] zeros
    uint4* sn;      // [SQR_A4][32] modulus limbs of every lane (NSM)
    uint4* ss;      // [SQR_X4][32] private stash: the slot accumulators rest here while a block product runs
    int wl;         // lane in the warp
    int gb;         // first lane of the group in the warp

    __device__ __forceinline__ void init_sqr(uint4* warp_smem) {
        wl = threadIdx.x & 31;
        gb = wl & ~(TPI - 1);
        sa = warp_smem;
        sx = warp_smem + SQR_A4 * 32;
        sz = warp_smem + (SQR_A4 + SQR_X4) * 32;
        sn = sz + 32;
        ss = sn + SQR_A4 * 32;
        sz[wl] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k4 = 0; k4 < L / 4; ++k4) sn[k4 * 32 + wl] = make_uint4(n[4 * k4], n[4 * k4 + 1], n[4 * k4 + 2], n[4 * k4 + 3]);
        __syncwarp();
    }

    // window (E, O, Ec, Oc) shifted down one limb, then += V * y.  Returns the limb the shift retires.
    __device__ __forceinline__ uint32_t prod_row(uint32_t (&E)[L], uint32_t (&O)[L], uint32_t& Ec, uint32_t& Oc,
                                                 const uint32_t (&V)[L], uint32_t y, uint32_t fence) {
        const uint32_t retired = E[0];
        uint32_t nE[L], nO[L], nEc, nOc, junk;
        // Row order fence: the carry-in below is always 0 (`fence` = lazy carry words of the row two back, < 2^31), but it
        // makes this row start only after that row's chains have finished: two rows = four carry chains are in flight,
        // enough to cover the carry latency.  Without a fence ptxas interleaves many rows, runs out of predicate
        // registers for their carries and packs them into bit masks (2-3 LOP3 per multiply).
        add_cc(junk, fence, fence);
        addc_cc(nE[0], O[0], E[1]);
#pragma unroll
        for (int j = 0; j < L - 2; j += 2) {
            madc_lo_cc(nO[j], V[j + 1], y, E[j + 2]);
            madc_hi_cc(nO[j + 1], V[j + 1], y, E[j + 3]);
        }
        madc_lo_cc(nO[L - 2], V[L - 1], y, Ec);
        madc_hi_cc(nO[L - 1], V[L - 1], y, 0);
        addc(nOc, 0, 0);
        mad_lo_cc(nE[0], V[0], y, nE[0]);
        madc_hi_cc(nE[1], V[0], y, O[1]);
#pragma unroll
        for (int j = 2; j < L; j += 2) {
            madc_lo_cc(nE[j], V[j], y, O[j]);
            madc_hi_cc(nE[j + 1], V[j], y, O[j + 1]);
        }
        addc(nEc, Oc, 0);
#pragma unroll
        for (int j = 0; j < L; ++j) { E[j] = nE[j]; O[j] = nO[j]; }
        Ec = nEc; Oc = nOc;
        return retired;
    }

    // final one-limb shift and merge of a window into plain limbs; returns the limb it retires
    __device__ __forceinline__ uint32_t window_close(uint32_t (&r)[L], uint32_t& ov, const uint32_t (&E)[L], const uint32_t (&O)[L],
                                                     uint32_t Ec, uint32_t Oc, uint32_t recv) {
        const uint64_t top = (uint64_t)Ec + recv;
        add_cc(r[0], O[0], E[1]);
#pragma unroll
        for (int k = 1; k < L - 1; ++k) addc_cc(r[k], O[k], E[k + 1]);
        addc_cc(r[L - 1], O[L - 1], (uint32_t)top);
        addc(ov, Oc, (uint32_t)(top >> 32));
        return E[0];
    }

    // buffer (2L limbs, uint4 rows) of this lane <- V[0..L) * y[0..NR), placed at limb offset `off` (multiple of 4);
    // the caller zeroes what the product does not cover.
    template <int NR>
    __device__ __forceinline__ void block_product_to_smem(const uint32_t (&V)[L], const uint4* __restrict__ yrow, int ysrc, int off4) {
        uint32_t E[L], O[L], Ec = 0, Oc = 0;
#pragma unroll
        for (int k = 0; k < L; ++k) { E[k] = 0; O[k] = 0; }
        uint32_t ret[4], fq[4] = {0, 0, 0, 0};   // lazy carry words of the last rows (row order fence)
#pragma unroll
        for (int j4 = 0; j4 < NR / 4; ++j4) {
            const uint4 y4 = yrow[j4 * 32 + ysrc];
            const uint32_t yy[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const uint32_t rl = prod_row(E, O, Ec, Oc, V, yy[jj], fq[PGPU_FENCE - 1]);
                fq[3] = fq[2]; fq[2] = fq[1]; fq[1] = fq[0]; fq[0] = Ec | Oc;
                // the row's shift retires the limb completed by the previous row
                if (jj > 0) ret[jj - 1] = rl;
                else if (j4 > 0) { ret[3] = rl; sx[(off4 + j4 - 1) * 32 + wl] = make_uint4(ret[0], ret[1], ret[2], ret[3]); }
            }
        }
        uint32_t hi[L], ov;
        ret[3] = window_close(hi, ov, E, O, Ec, Oc, 0u);
        sx[(off4 + NR / 4 - 1) * 32 + wl] = make_uint4(ret[0], ret[1], ret[2], ret[3]);
#pragma unroll
        for (int k4 = 0; k4 < L / 4; ++k4)
            sx[(off4 + NR / 4 + k4) * 32 + wl] = make_uint4(hi[4 * k4], hi[4 * k4 + 1], hi[4 * k4 + 2], hi[4 * k4 + 3]);
    }

    // acc (+ overflow word) += half `half` (0 low, 1 high) of lane `src`'s exchange buffer; stride4 = 32 rows apart,
    // or src row = zero row with stride4 = 0 for an absent term
    __device__ __forceinline__ void gather_add(uint32_t (&acc)[L], uint32_t& ov, const uint4* __restrict__ base, int stride4) {
        uint32_t v[L];
#pragma unroll
        for (int k4 = 0; k4 < L / 4; ++k4) {
            const uint4 q = base[k4 * stride4];
            v[4 * k4] = q.x; v[4 * k4 + 1] = q.y; v[4 * k4 + 2] = q.z; v[4 * k4 + 3] = q.w;
        }
        add_cc(acc[0], acc[0], v[0]);
#pragma unroll
        for (int k = 1; k < L; ++k) addc_cc(acc[k], acc[k], v[k]);
        addc(ov, ov, 0);
    }

    // the two slot accumulators leave the register file while a block product needs it
    __device__ __forceinline__ void stash(const uint32_t (&p)[L], const uint32_t (&q)[L]) {
#pragma unroll
        for (int k4 = 0; k4 < L / 4; ++k4) {
            ss[k4 * 32 + wl] = make_uint4(p[4 * k4], p[4 * k4 + 1], p[4 * k4 + 2], p[4 * k4 + 3]);
            ss[(L / 4 + k4) * 32 + wl] = make_uint4(q[4 * k4], q[4 * k4 + 1], q[4 * k4 + 2], q[4 * k4 + 3]);
        }
    }
    __device__ __forceinline__ void unstash(uint32_t (&p)[L], uint32_t (&q)[L]) {
#pragma unroll
        for (int k4 = 0; k4 < L / 4; ++k4) {
            const uint4 u = lds_v4_volatile(ss + k4 * 32 + wl), v = lds_v4_volatile(ss + (L / 4 + k4) * 32 + wl);
            p[4 * k4] = u.x; p[4 * k4 + 1] = u.y; p[4 * k4 + 2] = u.z; p[4 * k4 + 3] = u.w;
            q[4 * k4] = v.x; q[4 * k4 + 1] = v.y; q[4 * k4 + 2] = v.z; q[4 * k4 + 3] = v.w;
        }
    }

    __device__ __forceinline__ const uint4* xterm(int src_t, int half) const { return sx + (half * (L / 4)) * 32 + gb + src_t; }

    __device__ __forceinline__ void sqr(uint32_t (&r)[L], const uint32_t (&a)[L]) {
        if constexpr (!HAS_SQR) {
            mul(r, a, a);
        } else {
            // ---- operand to shared memory
#pragma unroll
            for (int k4 = 0; k4 < L / 4; ++k4) sa[k4 * 32 + wl] = make_uint4(a[4 * k4], a[4 * k4 + 1], a[4 * k4 + 2], a[4 * k4 + 3]);
            __syncwarp();
            const bool outer = (t == 0) || (t == 3);
            // The two slot accumulators P, Q live in the stash while a block product runs.  Round 0 is O_t followed by the
            // half product (P = heavy slot, Q = light slot, both doubled at its end and renamed to low/high slot); round 1
            // is D_t (P = low slot X^t, Q = high slot X^(t+4)).  One loop body serves both full-size products, which keeps
            // the code of the whole squaring within reach of the instruction cache.
            uint32_t p_ov = 0, q_ov = 0;
            {
                uint32_t Z[L];
#pragma unroll
                for (int k = 0; k < L; ++k) Z[k] = 0;
                stash(Z, Z);
            }
#pragma unroll 1
            for (int round = 0; round < 2; ++round) {
                // O_t = A_t * A_(t+1 mod 4), base slot t + (t+1 mod 4)   |   D_t = A_t^2, base slot 2t
                block_product_to_smem<L>(a, sa, round == 0 ? gb + ((t + 1) & 3) : wl, 0);
                __syncwarp();
                uint32_t P[L], Q[L];
                unstash(P, Q);
                if (round == 0) {
                    const int s1 = (t == 0) ? 1 : (t == 1) ? 2 : (t == 2) ? 0 : 1;
                    const int h1 = (t == 0 || t == 2) ? 1 : 0;
                    gather_add(P, p_ov, xterm(s1, h1), 32);
                    // second heavy term for lanes 0 and 3 (from lane 3), light term for lanes 1 and 2
                    gather_add(P, p_ov, outer ? xterm(3, t == 0 ? 1 : 0) : sz + wl, outer ? 32 : 0);
                    gather_add(Q, q_ov, outer ? sz + wl : xterm(t == 1 ? 0 : 2, t == 1 ? 0 : 1), outer ? 0 : 32);
                    stash(P, Q);
                    __syncwarp();
                    // half product: lanes 0,2 share A_0 * A_2, lanes 1,3 share A_1 * A_3; base slot 2 + 2*(t&1)
                    {
                        uint32_t VH[L];
#pragma unroll
                        for (int k4 = 0; k4 < L / 4; ++k4) {
                            const uint4 q = sa[k4 * 32 + gb + (t & 1)];
                            VH[4 * k4] = q.x; VH[4 * k4 + 1] = q.y; VH[4 * k4 + 2] = q.z; VH[4 * k4 + 3] = q.w;
                        }
                        const int hoff4 = (t < 2) ? 0 : L / 8;             // product placed at limb offset 0 or L/2
                        const int zoff4 = (t < 2) ? 3 * L / 8 : 0;         // zero padding on the other side
#pragma unroll
                        for (int k4 = 0; k4 < L / 8; ++k4) sx[(zoff4 + k4) * 32 + wl] = make_uint4(0, 0, 0, 0);
                        block_product_to_smem<L / 2>(VH, sa + hoff4 * 32, gb + (t & 1) + 2, hoff4);
                    }
                    __syncwarp();
                    unstash(P, Q);
                    {
                        const int p = (t < 2) ? 1 : 0;
                        gather_add(P, p_ov, xterm(p, t & 1), 32);
                        gather_add(P, p_ov, xterm(p + 2, t & 1), 32);
                    }
                    // the cross terms count twice
                    p_ov = (p_ov << 1) | (P[L - 1] >> 31);
                    q_ov = (q_ov << 1) | (Q[L - 1] >> 31);
#pragma unroll
                    for (int k = L - 1; k > 0; --k) {
                        P[k] = __funnelshift_l(P[k - 1], P[k], 1);
                        Q[k] = __funnelshift_l(Q[k - 1], Q[k], 1);
                    }
                    P[0] <<= 1; Q[0] <<= 1;
                    // heavy/light -> low/high slot
                    {
                        const bool swap = t < 2;
#pragma unroll
                        for (int k = 0; k < L; ++k) { const uint32_t x = P[k], y = Q[k]; P[k] = swap ? y : x; Q[k] = swap ? x : y; }
                        const uint32_t x = p_ov, y = q_ov;
                        p_ov = swap ? y : x; q_ov = swap ? x : y;
                    }
                } else {
                    gather_add(P, p_ov, xterm(t >> 1, t & 1), 32);
                    gather_add(Q, q_ov, xterm(2 + (t >> 1), t & 1), 32);
                }
                stash(P, Q);
                __syncwarp();
            }
            uint32_t lo[L], hi[L];
            unstash(lo, hi);
            const uint32_t lo_ov = p_ov, hi_ov = q_ov;
            // ---- phase 2: Montgomery reduction of the low half (lane t: lo + lo_ov * X), sliding window as in mul()
            uint32_t E[L], O[L], Ec = 0, Oc = lo_ov, nn[L];
            fetch_n(nn);
#pragma unroll
            for (int k = 0; k < L; ++k) { E[k] = 0; O[k] = lo[k]; }
#pragma unroll 1
            for (int u = 0; u < TPI; ++u) {
#pragma unroll
                for (int k = 0; k < L; k += 2) {
                    // two quotient digits at once (radix 2^64): both rows' chains can then run side by side
                    const uint64_t s64 = ((uint64_t)E[1] | ((uint64_t)E[2] << 32)) + ((uint64_t)O[0] | ((uint64_t)O[1] << 32));
                    const uint64_t q64 = s64 * np64;
                    const uint32_t qq[2] = {__shfl_sync(FULL_MASK, (uint32_t)q64, 0, TPI), __shfl_sync(FULL_MASK, (uint32_t)(q64 >> 32), 0, TPI)};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t q = qq[h];
                        uint32_t recv = __shfl_down_sync(FULL_MASK, E[0], 1, TPI);
                        if (t == TPI - 1) recv = 0;
                        const uint64_t top = (uint64_t)Ec + recv;
                        const uint32_t top_lo = (uint32_t)top, top_hi = (uint32_t)(top >> 32);
                        uint32_t nE[L], nO[L], nEc, nOc;
                        add_cc(nE[0], O[0], E[1]);
                        // n[odd] * q onto the old even array >> 64
#pragma unroll
                        for (int j = 0; j < L - 2; j += 2) {
                            madc_lo_cc(nO[j], nn[j + 1], q, E[j + 2]);
                            madc_hi_cc(nO[j + 1], nn[j + 1], q, E[j + 3]);
                        }
                        madc_lo_cc(nO[L - 2], nn[L - 1], q, top_lo);
                        madc_hi_cc(nO[L - 1], nn[L - 1], q, top_hi);
                        addc(nOc, 0, 0);
                        // n[even] * q onto the old odd array
                        mad_lo_cc(nE[0], nn[0], q, nE[0]);
                        madc_hi_cc(nE[1], nn[0], q, O[1]);
#pragma unroll
                        for (int j = 2; j < L; j += 2) {
                            madc_lo_cc(nE[j], nn[j], q, O[j]);
                            madc_hi_cc(nE[j + 1], nn[j], q, O[j + 1]);
                        }
                        addc(nEc, Oc, 0);
#pragma unroll
                        for (int j = 0; j < L; ++j) { E[j] = nE[j]; O[j] = nO[j]; }
                        Ec = nEc; Oc = nOc;
                    }
                }
            }
            uint32_t recv = __shfl_down_sync(FULL_MASK, E[0], 1, TPI);
            if (t == TPI - 1) recv = 0;
            uint32_t ov;
            window_close(r, ov, E, O, Ec, Oc, recv);
            // ---- + high half
            uint32_t cy;
            add_cc(r[0], r[0], hi[0]);
#pragma unroll
            for (int k = 1; k < L; ++k) addc_cc(r[k], r[k], hi[k]);
            addc(cy, 0, 0);
            resolve_reduce(r, ov + hi_ov + cy);
        }
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
        uint32_t s[L], cy, nn[L];
        fetch_n(nn);
        add_cc(s[0], d[0], nn[0]);
#pragma unroll
        for (int k = 1; k < L; ++k) addc_cc(s[k], d[k], nn[k]);
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
