// Host-side launch interface of the kernels in powm.cu / aux_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "vm.h"

namespace pgpu {

constexpr int VM_BLOCK_THREADS = 128;

// A built instantiation of the exponentiation kernel for moduli of up to S 32-bit limbs: TPI lanes per residue,
// L limbs per lane.  fp64 = false: 32-bit limbs on the integer pipe (powm_vm<tpi, L>, S = tpi*L, R = 2^(32*S));
// fp64 = true: 52-bit limbs on the FP64 pipe (powm_vm52<tpi, L, S>, R = 2^(52*tpi*L)).
struct VmShape {
    int S, tpi, L;
    bool fp64 = false;
    int rbits() const { return fp64 ? 52 * tpi * L : 32 * S; }          // Montgomery radix R = 2^rbits
    int tbl_limbs() const { return fp64 ? 2 * tpi * L : S; }            // 32-bit words of one scratch-table entry
};

// Launch the kernel of `sh` on `blocks` blocks of VM_BLOCK_THREADS threads.
cudaError_t vm_launch(const VmShape& sh, const VmParams& P, int blocks, cudaStream_t stream);
// Resident blocks per SM for that instantiation (0 if the shape is not built).
int vm_occupancy(const VmShape& sh);

}  // namespace pgpu
