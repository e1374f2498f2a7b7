// Launch interface of aux_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pgpu {

constexpr int CRT_MAXH = 64;   // limbs of p, q supported by the CRT recombination (4096-bit n)

// consts: 7 records of h limbs: p, q, p^-1 mod 2^(32h), q^-1 mod 2^(32h),
// h_p*2^(32h) mod p, h_q*2^(32h) mod q, q^-1*2^(32h) mod p
struct CrtParams {
    uint32_t n_items;
    int h;
    const uint32_t* consts;
    uint32_t np0_p, np0_q;
    const uint32_t* xp;     // c^(p-1) mod p^2, x_stride limbs apart
    const uint32_t* xq;     // c^(q-1) mod q^2
    uint32_t x_stride;
    uint32_t* out;          // plaintexts, out_stride limbs apart, out_limbs written
    uint32_t out_stride;
    uint32_t out_limbs;
};

cudaError_t crt_combine_launch(const CrtParams& P, cudaStream_t stream);

// out (S limbs) = product of n_items records (stride S) times 2^(-32*S*T) mod N,
// T = number of Montgomery multiplications performed (returned by prod_reduce_mults()).
struct ProdParams {
    const uint32_t* in;      // n_items records of S limbs
    uint32_t n_items;
    const uint32_t* mod;
    uint32_t np0;
    uint32_t* partial;       // one record per block
};
cudaError_t prod_reduce_launch(int tpi, int limbs, const ProdParams& P, int blocks, cudaStream_t stream);

}  // namespace pgpu
