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
#include "vm_run.cuh"
#include "launch.h"

namespace pgpu {

template <int TPI, int L>
__global__ void __launch_bounds__(VM_BLOCK_THREADS, (L <= 16 ? 4 : L <= 32 ? 2 : 1)) powm_vm(const VmParams P) {
    vm_run<Vm32<TPI, L>>(P);
}

// resident blocks per SM the FP64 shapes are compiled for (register budget 65536 / (128 * blocks))
constexpr int vm52_min_blocks(int L) { return L <= 5 ? 6 : L <= 8 ? 4 : L <= 10 ? 3 : 2; }

template <int TPI, int L, int S32>
__global__ void __launch_bounds__(VM_BLOCK_THREADS, vm52_min_blocks(L)) powm_vm52(const VmParams P) {
    vm_run<Vm52<TPI, L, S32>>(P);
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

template <int TPI, int L, int S32>
static cudaError_t launch52_t(const VmParams& P, int blocks, cudaStream_t stream) {
    powm_vm52<TPI, L, S32><<<blocks, VM_BLOCK_THREADS, 0, stream>>>(P);
    return cudaGetLastError();
}

template <int TPI, int L, int S32>
static int occupancy52_t() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, powm_vm52<TPI, L, S32>, VM_BLOCK_THREADS, 0);
    return nb;
}

#define PGPU_FOR_EACH_SHAPE(X) \
    X(2, 16) X(4, 8)           \
    X(4, 16) X(8, 8)           \
    X(8, 12) X(4, 24)          \
    X(8, 16) X(16, 8) X(4, 32) X(32, 4) \
    X(8, 24) X(16, 12) X(32, 6)

// FP64-pipe shapes (TPI, L, 32-bit record limbs): 52*TPI*L >= 32*S32 + 2
#define PGPU_FOR_EACH_SHAPE52(X) \
    X(4, 5, 32)                  \
    X(4, 10, 64) X(8, 5, 64)     \
    X(4, 15, 96) X(8, 8, 96)     \
    X(8, 10, 128)                \
    X(8, 15, 192) X(16, 8, 192)

cudaError_t vm_launch(const VmShape& sh, const VmParams& P, int blocks, cudaStream_t stream) {
    if (sh.fp64) {
#define X(T, LL, SS) if (sh.tpi == T && sh.L == LL && sh.S == SS) return launch52_t<T, LL, SS>(P, blocks, stream);
        PGPU_FOR_EACH_SHAPE52(X)
#undef X
        return cudaErrorInvalidValue;
    }
#define X(T, LL) if (sh.tpi == T && sh.L == LL && sh.S == T * LL) return launch_t<T, LL>(P, blocks, stream);
    PGPU_FOR_EACH_SHAPE(X)
#undef X
    return cudaErrorInvalidValue;
}

int vm_occupancy(const VmShape& sh) {
    if (sh.fp64) {
#define X(T, LL, SS) if (sh.tpi == T && sh.L == LL && sh.S == SS) return occupancy52_t<T, LL, SS>();
        PGPU_FOR_EACH_SHAPE52(X)
#undef X
        return 0;
    }
#define X(T, LL) if (sh.tpi == T && sh.L == LL && sh.S == T * LL) return occupancy_t<T, LL>();
    PGPU_FOR_EACH_SHAPE(X)
#undef X
    return 0;
}

}  // namespace pgpu
