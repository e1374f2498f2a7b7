// Host-side launch interface of the kernels in powm.cu / aux_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "vm.h"

namespace pgpu {

constexpr int VM_BLOCK_THREADS = 128;

// Launch powm_vm<tpi, limbs> on `blocks` blocks of VM_BLOCK_THREADS threads.
cudaError_t vm_launch(int tpi, int limbs, const VmParams& P, int blocks, cudaStream_t stream);
// Resident blocks per SM for that instantiation (0 if the shape is not built).
int vm_occupancy(int tpi, int limbs);

}  // namespace pgpu
