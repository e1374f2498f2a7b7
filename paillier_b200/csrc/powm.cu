// powm_vm: the batched modular-exponentiation engine (see vm.h, mont.cuh).
//
// Persistent grid: gridDim.x * blockDim.x / TPI groups are resident; group g
// handles items g, g + n_groups, ...
// The group's window table lives in a private slot of a global scratch buffer
// ([entry][group][S] so that neighbouring groups coalesce); at 16 entries x
// 512 B it stays in the 126 MB L2 for every resident group.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "mont.cuh"
#include "vm.h"
#include "launch.h"

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

template <int TPI, int L>
__global__ void __launch_bounds__(VM_BLOCK_THREADS, (L <= 16 ? 4 : L <= 32 ? 2 : 1)) powm_vm(const VmParams P) {
    constexpr int S = TPI * L;
    using MontT = Mont<TPI, L, SqrShape<TPI, L>::value>;
    MontT M;
    const uint32_t n_groups = P.n_groups;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / TPI;
    const uint32_t n_inst = P.n_items;
    const uint32_t rounds = (n_inst + n_groups - 1) / n_groups;
    const int lane_t = (threadIdx.x & 31) & (TPI - 1);
    uint32_t* const tbl = P.table + (size_t)group * S + lane_t * L;
    const size_t tbl_entry_stride = (size_t)n_groups * S;
    M.init(P.mod, P.np0);
    if constexpr (MontT::HAS_SQR) {
        extern __shared__ uint4 vm_smem[];
        M.init_sqr(vm_smem + (threadIdx.x >> 5) * (MontT::SQR_ROWS * 32));
    }
    const uint32_t* const kc = P.kconst + lane_t * L;
    uint32_t* const dump = P.dump + (size_t)group * S;

    for (uint32_t rd = 0; rd < rounds; ++rd) {
        uint32_t item = rd * n_groups + group;
        const bool active = item < n_inst;      // whole warps stay in lock-step; idle groups redo item 0
        if (!active) item = 0;

        uint32_t x[L], y[L], y2[L];
#pragma unroll
        for (int k = 0; k < L; ++k) x[k] = 0;

        for (const uint32_t* pc = P.prog;; ++pc) {
            const uint32_t op = __ldg(pc);
            const uint32_t code = op >> 27, arg = op & 0x07ffffffu;
            if (code == OP_END) break;
            uint32_t nsq = 0, nmul = 0, bkt = 0xffffffffu;
            switch (code) {
                case OP_LDI: {
                    const uint32_t* p = P.in[arg] + (size_t)(item / P.in_div[arg]) * P.in_stride[arg];
                    const uint32_t lim = P.in_limbs[arg];
#pragma unroll
                    for (int k = 0; k < L; ++k) {
                        const uint32_t idx = lane_t * L + k;
                        x[k] = idx < lim ? __ldg(p + idx) : 0u;
                    }
                } break;
                case OP_LDC: load_vec<L>(x, kc + (size_t)arg * S); break;
                case OP_LDT: load_vec<L>(x, tbl + arg * tbl_entry_stride); break;
                case OP_STT: store_vec<L>(tbl + arg * tbl_entry_stride, x); break;
                case OP_STO: {
                    // idle groups (they redo item 0 in lock step) store into their dump record: no branch on `active`,
                    // which would make the compiler clone the whole interpreter loop
                    uint32_t* p = active ? P.out[arg] + (size_t)item * P.out_stride[arg] : dump;
                    const uint32_t lim = P.out_limbs[arg];
#pragma unroll
                    for (int k = 0; k < L; ++k) {
                        const uint32_t idx = lane_t * L + k;
                        if (idx < lim) p[idx] = x[k];
                    }
                } break;
                case OP_SQR: nsq = arg; break;
                case OP_MULT: load_vec<L>(y, tbl + arg * tbl_entry_stride); nmul = 1; break;
                case OP_MULC: load_vec<L>(y, kc + (size_t)arg * S); nmul = 1; break;
                case OP_MULI: {
                    const uint32_t* p = P.in[arg] + (size_t)(item / P.in_div[arg]) * P.in_stride[arg];
                    const uint32_t lim = P.in_limbs[arg];
#pragma unroll
                    for (int k = 0; k < L; ++k) {
                        const uint32_t idx = lane_t * L + k;
                        y[k] = idx < lim ? __ldg(p + idx) : 0u;
                    }
                    nmul = 1;
                } break;
                case OP_ADDT: load_vec<L>(y, tbl + arg * tbl_entry_stride); M.add(x, x, y); break;
                case OP_ADDC: load_vec<L>(y, kc + (size_t)arg * S); M.add(x, x, y); break;
                case OP_WIN: {
                    const uint32_t pos = arg & 0xfffffu, w = (arg >> 20) & 0xfu, tbase = arg >> 24;
                    const uint32_t idx = exp_bits(P.exp + (size_t)item * P.exp_stride, P.exp_bits, pos, w);
                    load_vec<L>(y, tbl + (tbase + idx) * tbl_entry_stride);
                    nsq = w; nmul = 1;
                } break;
                case OP_FIXW: {
                    const uint32_t pos = arg & 0xfffffu, w = (arg >> 20) & 0xfu;
                    const uint32_t idx = exp_bits(P.exp + (size_t)item * P.exp_stride, P.exp_bits, pos, w);
                    load_vec<L>(y, P.fixed + ((size_t)(pos / w) * (1u << w) + idx) * S + lane_t * L);
                    nmul = 1;
                } break;
                case OP_LDIO:
                    load_vec<L>(x, P.in[arg & 3u] + (size_t)item * P.in_stride[arg & 3u] + (size_t)(arg >> 2) * S + lane_t * L);
                    break;
                case OP_MULIO:
                    load_vec<L>(y, P.in[arg & 3u] + (size_t)item * P.in_stride[arg & 3u] + (size_t)(arg >> 2) * S + lane_t * L);
                    nmul = 1;
                    break;
                case OP_STOO: {
                    const uint32_t a = arg & 1u, off = arg >> 2;
                    uint32_t* p = active ? P.out[a] + (size_t)item * P.out_stride[a] + (size_t)off * S : dump;
                    store_vec<L>(p + lane_t * L, x);
                } break;
                case OP_BKT: {
                    const uint32_t pos = arg & 0xfffffu, w = (arg >> 20) & 0xfu;
                    // idle groups must not touch the buckets: they multiply into the spare entry 2^w
                    bkt = active ? exp_bits(P.exp + (size_t)item * P.exp_stride, P.exp_bits, pos, w) : (1u << w);
                    load_vec<L>(y, tbl + bkt * tbl_entry_stride);
                    nmul = 1;
                } break;
                case OP_SUBT: load_vec<L>(y, tbl + arg * tbl_entry_stride); M.sub(x, x, y); break;
                case OP_SQMT: {
                    const uint32_t idx = arg >> 12;
                    load_vec<L>(y, tbl + idx * tbl_entry_stride);
                    nsq = arg & 0xfffu; nmul = 1;
                } break;
                default: break;
            }
            // the single Montgomery multiplier of the instruction stream
            if constexpr (MontT::HAS_SQR) {
                if (P.flags & 1u) {
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
                    uint32_t b[L];
#pragma unroll
                    for (int k = 0; k < L; ++k) b[k] = sq ? x[k] : y[k];
                    M.mul(x, x, b);
                }
            }
            if (bkt != 0xffffffffu) store_vec<L>(tbl + bkt * tbl_entry_stride, x);
        }
    }
}

template <int TPI, int L>
constexpr size_t vm_smem_bytes() {
    using MontT = Mont<TPI, L, SqrShape<TPI, L>::value>;
    return MontT::HAS_SQR ? MontT::SQR_SMEM_PER_WARP * (VM_BLOCK_THREADS / 32) : 0;
}

template <int TPI, int L>
static cudaError_t prepare_t() {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && done[dev]) return cudaSuccess;
    if (vm_smem_bytes<TPI, L>() > 48 * 1024) {
        e = cudaFuncSetAttribute(powm_vm<TPI, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vm_smem_bytes<TPI, L>());
        if (e != cudaSuccess) return e;
    }
    if (dev < 64) done[dev] = true;
    return cudaSuccess;
}

template <int TPI, int L>
static cudaError_t launch_t(const VmParams& P, int blocks, cudaStream_t stream) {
    cudaError_t e = prepare_t<TPI, L>();
    if (e != cudaSuccess) return e;
    powm_vm<TPI, L><<<blocks, VM_BLOCK_THREADS, vm_smem_bytes<TPI, L>(), stream>>>(P);
    return cudaGetLastError();
}

template <int TPI, int L>
static int occupancy_t() {
    int nb = 0;
    if (prepare_t<TPI, L>() != cudaSuccess) return 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, powm_vm<TPI, L>, VM_BLOCK_THREADS, vm_smem_bytes<TPI, L>());
    return nb;
}

#define PGPU_FOR_EACH_SHAPE(X) \
    X(2, 16) X(4, 8)           \
    X(4, 16) X(8, 8)           \
    X(8, 12) X(4, 24)          \
    X(8, 16) X(16, 8) X(4, 32) X(32, 4) \
    X(8, 24) X(16, 12) X(32, 6)

cudaError_t vm_launch(int tpi, int limbs, const VmParams& P, int blocks, cudaStream_t stream) {
#define X(T, LL) if (tpi == T && limbs == LL) return launch_t<T, LL>(P, blocks, stream);
    PGPU_FOR_EACH_SHAPE(X)
#undef X
    return cudaErrorInvalidValue;
}

int vm_occupancy(int tpi, int limbs) {
#define X(T, LL) if (tpi == T && limbs == LL) return occupancy_t<T, LL>();
    PGPU_FOR_EACH_SHAPE(X)
#undef X
    return 0;
}

}  // namespace pgpu
