// vm_run: the interpreter of the micro-programs (vm.h) behind powm_vm / powm_vm52 (powm.cu), with the two multipliers it can
// run on.  A header of its own so that tests/cpp/mont_host_test.cpp can execute the SAME source on an emulated warp of CPU
// threads (tests/cpp/cuda_host_shim.h) against programs the library compiled; powm.cu wraps it in the __global__ kernels.
#pragma once
#include <cstdint>
#include "mont.cuh"
#include "mont52.cuh"
#include "vm.h"

namespace pgpu {

template <int L>
__device__ __forceinline__ void load_vec(uint32_t (&x)[L], const uint32_t* __restrict__ p) {
    if constexpr (L % 4 == 0) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int k = 0; k < L / 4; ++k) {
            uint4 v = q[k];
            x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
        }
    } else {
        const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
        for (int k = 0; k < L / 2; ++k) {
            uint2 v = q[k];
            x[2 * k] = v.x; x[2 * k + 1] = v.y;
        }
    }
}

template <int L>
__device__ __forceinline__ void store_vec(uint32_t* __restrict__ p, const uint32_t (&x)[L]) {
    if constexpr (L % 4 == 0) {
        uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
        for (int k = 0; k < L / 4; ++k) q[k] = make_uint4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    } else {
        uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
        for (int k = 0; k < L / 2; ++k) q[k] = make_uint2(x[2 * k], x[2 * k + 1]);
    }
}


// ---- the two multipliers behind one interpreter -----------------------------------------------------------------
// Vm32: 32-bit limbs on the integer pipe (mont.cuh); registers, table entries and records share one format.
// Mont52 (mont52.cuh): 52-bit limbs as doubles on the FP64 pipe; records are converted when they are loaded or stored.
// Pointers handed to load_rec / store_rec / load_full are record starts, to load_tbl / store_tbl the group's slot.
template <int TPI, int L>
struct Vm32 : Mont<TPI, L, SqrShape<TPI, L>::value> {
    using Base = Mont<TPI, L, SqrShape<TPI, L>::value>;
    using elem = uint32_t;
    static constexpr int TPI_ = TPI, L_ = L;
    static constexpr int S32 = TPI * L;
    static constexpr int TBL = L;

    __device__ __forceinline__ void load_full(uint32_t (&x)[L], const uint32_t* __restrict__ p) const { load_vec<L>(x, p + this->t * L); }
    __device__ __forceinline__ void store_full(uint32_t* __restrict__ p, const uint32_t (&x)[L]) const { store_vec<L>(p + this->t * L, x); }
    // records narrower than the modulus read as zero-extended; full records stream in 128-bit vectors
    __device__ __forceinline__ void load_rec(uint32_t (&x)[L], const uint32_t* __restrict__ p, uint32_t lim) const {
        if (lim == (uint32_t)S32 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) { load_full(x, p); return; }
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const uint32_t idx = this->t * L + k;
            x[k] = idx < lim ? __ldg(p + idx) : 0u;
        }
    }
    __device__ __forceinline__ void store_rec(uint32_t* __restrict__ p, const uint32_t (&x)[L], uint32_t lim) const {
        if (lim == (uint32_t)S32 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) { store_full(p, x); return; }
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const uint32_t idx = this->t * L + k;
            if (idx < lim) p[idx] = x[k];
        }
    }
    __device__ __forceinline__ void load_tbl(uint32_t (&x)[L], const uint32_t* __restrict__ p) const { load_vec<L>(x, p + this->t * L); }
    __device__ __forceinline__ void store_tbl(uint32_t* __restrict__ p, const uint32_t (&x)[L]) const { store_vec<L>(p + this->t * L, x); }
};

template <int TPI, int L, int S32_>
struct Vm52 : Mont52<TPI, L, S32_> {
    using Base = Mont52<TPI, L, S32_>;
    static constexpr bool HAS_SQR = false;
    static constexpr int TPI_ = TPI, L_ = L;
    __device__ __forceinline__ void load_full(double (&x)[L], const uint32_t* __restrict__ p) const { this->load_rec(x, p, S32_); }
    __device__ __forceinline__ void store_full(uint32_t* __restrict__ p, const double (&x)[L]) const { this->store_rec(p, x, S32_); }
    __device__ __forceinline__ void load_tbl(double (&x)[L], const uint32_t* __restrict__ p) const { Base::load_tbl(x, p + this->t * Base::TBL); }
    __device__ __forceinline__ void store_tbl(uint32_t* __restrict__ p, const double (&x)[L]) const { Base::store_tbl(p + this->t * Base::TBL, x); }
};

// bits [pos, pos+w) of a little-endian limb array
__device__ __forceinline__ uint32_t exp_bits(const uint32_t* __restrict__ e, uint32_t nbits, uint32_t pos, uint32_t w) {
    if (pos >= nbits) return 0;
    const uint32_t limb = pos >> 5, sh = pos & 31, nlimbs = (nbits + 31) >> 5;
    uint64_t v = e[limb];
    if (sh + w > 32 && limb + 1 < nlimbs) v |= (uint64_t)e[limb + 1] << 32;
    uint32_t r = (uint32_t)(v >> sh) & ((1u << w) - 1u);
    if (pos + w > nbits) r &= (1u << (nbits - pos)) - 1u;
    return r;
}

template <class B>
__device__ __forceinline__ void vm_run(const VmParams& P) {
    constexpr int L = B::L_;
    constexpr int TPI = B::TPI_;
    constexpr int S = B::S32;                  // 32-bit limbs of a record
    constexpr int TS = TPI * B::TBL;           // 32-bit words of one table entry of a group
    using elem = typename B::elem;
    B M;
    const uint32_t n_groups = P.n_groups;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / TPI;
    const uint32_t n_inst = P.n_items;
    const uint32_t rounds = (n_inst + n_groups - 1) / n_groups;
    uint32_t* const tbl = P.table + (size_t)group * TS;
    const size_t tbl_entry_stride = (size_t)n_groups * TS;
    M.init(P.mod, P.np0);
    if constexpr (B::HAS_SQR) {
#ifndef PGPU_HOST_EMULATION
        extern __shared__ uint4 vm_smem[];
#else
        uint4* const vm_smem = hostwarp::shared_mem;       // the emulated block's shared memory (cuda_host_shim.h)
#endif
        M.init_sqr(vm_smem + (threadIdx.x >> 5) * (B::SQR_ROWS * 32));
    }
    const uint32_t* const kc = P.kconst;
    uint32_t* const dump = P.dump + (size_t)group * S;

    for (uint32_t rd = 0; rd < rounds; ++rd) {
        uint32_t item = rd * n_groups + group;
        const bool active = item < n_inst;      // whole warps stay in lock-step; idle groups redo item 0
        if (!active) item = 0;

        elem x[L], y[L];
#pragma unroll
        for (int k = 0; k < L; ++k) x[k] = 0;

        for (const uint32_t* pc = P.prog;; ++pc) {
            const uint32_t op = __ldg(pc);
            const uint32_t code = op >> 27, arg = op & 0x07ffffffu;
            if (code == OP_END) break;
            uint32_t nsq = 0, nmul = 0, bkt = 0xffffffffu;
            switch (code) {
                case OP_LDI: M.load_rec(x, P.in[arg] + (size_t)(item / P.in_div[arg]) * P.in_stride[arg], P.in_limbs[arg]); break;
                case OP_LDC: M.load_full(x, kc + (size_t)arg * S); break;
                case OP_LDT: M.load_tbl(x, tbl + arg * tbl_entry_stride); break;
                case OP_STT: M.store_tbl(tbl + arg * tbl_entry_stride, x); break;
                case OP_STO:
                    // idle groups (they redo item 0 in lock step) store into their dump record: no branch on `active`,
                    // which would make the compiler clone the whole interpreter loop
                    M.store_rec(active ? P.out[arg] + (size_t)item * P.out_stride[arg] : dump, x, P.out_limbs[arg]);
                    break;
                case OP_SQR: nsq = arg; break;
                case OP_MULT: M.load_tbl(y, tbl + arg * tbl_entry_stride); nmul = 1; break;
                case OP_MULC: M.load_full(y, kc + (size_t)arg * S); nmul = 1; break;
                case OP_MULI:
                    M.load_rec(y, P.in[arg] + (size_t)(item / P.in_div[arg]) * P.in_stride[arg], P.in_limbs[arg]);
                    nmul = 1;
                    break;
                case OP_ADDT: M.load_tbl(y, tbl + arg * tbl_entry_stride); M.add(x, x, y); break;
                case OP_ADDC: M.load_full(y, kc + (size_t)arg * S); M.add(x, x, y); break;
                case OP_WIN: {
                    const uint32_t pos = arg & 0xfffffu, w = (arg >> 20) & 0xfu, tbase = arg >> 24;
                    const uint32_t idx = exp_bits(P.exp + (size_t)item * P.exp_stride, P.exp_bits, pos, w);
                    M.load_tbl(y, tbl + (tbase + idx) * tbl_entry_stride);
                    nsq = w; nmul = 1;
                } break;
                case OP_FIXW: {
                    const uint32_t pos = arg & 0xfffffu, w = (arg >> 20) & 0xfu;
                    const uint32_t idx = exp_bits(P.exp + (size_t)item * P.exp_stride, P.exp_bits, pos, w);
                    M.load_full(y, P.fixed + ((size_t)(pos / w) * (1u << w) + idx) * S);
                    nmul = 1;
                } break;
                case OP_LDIO:
                    M.load_full(x, P.in[arg & 3u] + (size_t)item * P.in_stride[arg & 3u] + (size_t)(arg >> 2) * S);
                    break;
                case OP_MULIO:
                    M.load_full(y, P.in[arg & 3u] + (size_t)item * P.in_stride[arg & 3u] + (size_t)(arg >> 2) * S);
                    nmul = 1;
                    break;
                case OP_STOO: {
                    const uint32_t a = arg & 1u, off = arg >> 2;
                    M.store_full(active ? P.out[a] + (size_t)item * P.out_stride[a] + (size_t)off * S : dump, x);
                } break;
                case OP_BKT: {
                    const uint32_t pos = arg & 0xfffffu, w = (arg >> 20) & 0xfu, sub = arg >> 24;
                    // idle groups must not touch the buckets: they multiply into the spare entry 2^w
                    const uint32_t d = active ? exp_bits(P.exp + (size_t)item * P.exp_stride + (size_t)sub * P.exp_sub, P.exp_bits, pos, w) : (1u << w);
                    bkt = sub * ((1u << w) + 1u) + d;
                    M.load_tbl(y, tbl + bkt * tbl_entry_stride);
                    nmul = 1;
                } break;
                case OP_SUBT: M.load_tbl(y, tbl + arg * tbl_entry_stride); M.sub(x, x, y); break;
                case OP_SQMT: {
                    const uint32_t idx = arg >> 12;
                    M.load_tbl(y, tbl + idx * tbl_entry_stride);
                    nsq = arg & 0xfffu; nmul = 1;
                } break;
                default: break;
            }
            // the single Montgomery multiplier of the instruction stream
            if constexpr (B::HAS_SQR) {
                if (P.flags & 1u) {
                    elem y2[L];
#pragma unroll 1
                    for (uint32_t i = nsq; i > 0; --i) {
#pragma unroll
                        for (int k = 0; k < L; ++k) y2[k] = x[k];
                        M.mul(x, x, y2);
                    }
                } else {
#pragma unroll 1
                    for (uint32_t i = nsq; i > 0; --i) M.sqr(x, x);
                }
                if (nmul) M.mul(x, x, y);
            } else {
                for (uint32_t i = nsq + nmul; i > 0; --i) {
                    const bool sq = i > nmul;
                    elem b[L];
#pragma unroll
                    for (int k = 0; k < L; ++k) b[k] = sq ? x[k] : y[k];
                    M.mul(x, x, b);
                }
            }
            if (bkt != 0xffffffffu) M.store_tbl(tbl + bkt * tbl_entry_stride, x);
        }
    }
}

}  // namespace pgpu
