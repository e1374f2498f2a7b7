// FP64-pipe microbenchmark for B200 (sm_100a): could a 52-bit-limb Montgomery multiplier on DFMA beat the
// IMAD.WIDE one?  ("measure before building".)
//
// A 52x52-bit limb product costs, in the double-precision scheme (two DFMA.RZ extract the high and low
// 52 bits, one DADD in between, two 64-bit integer accumulations on the bit patterns):
//     hi = fma_rz(a, b, 2^104);  s = (2^104 + 2^52) - hi;  lo = fma_rz(a, b, s);  acc_hi += bits(hi);  acc_lo += bits(lo)
// i.e. 3 FP64-pipe instructions + 2 IADD3/IADD3.X pairs (3-input adds fold two accumulations into one pair)
// for 2704 bit^2, against one IMAD.WIDE.U32 for 1024 bit^2.  This tool measures
//   (1) the dependency-free DFMA issue rate,
//   (2) the rate of the whole limb-product pattern above,
//   (3) DFMA and IMAD.WIDE issued side by side (do the two pipes overlap?),
// and prints bit^2/s next to the IMAD.WIDE figure of tools/imad_peak.cu.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_peak dfma_peak.cu
// Run:   ./dfma_peak [json-out]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ACC = 8;
constexpr int INNER = 32;

__global__ void k_dfma_pure(uint64_t* out, double a0, double b0, int trips) {
    double acc[ACC], a[ACC], b[4];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i] = i + threadIdx.x; a[i] = a0 * (i + 3) + threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 * (2 * i + 1) + blockIdx.x;
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < INNER; ++r) {
#pragma unroll
            for (int i = 0; i < ACC; ++i) acc[i] = __fma_rz(a[i], b[r & 3], acc[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    if (s == 1.2345) out[0] = (uint64_t)s;
}

// the limb-product pattern: ACC products per repetition, each 2 DFMA + 1 DADD + two 64-bit integer accumulations
__global__ void k_dfma_product(uint64_t* out, double a0, double b0, int trips) {
    const double C1 = 0x1p104, C2 = 0x1.0000000000001p104;   // 2^104, 2^104 + 2^52
    double a[ACC], b[4];
    uint64_t lo[ACC], hi[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { a[i] = a0 * (i + 3) + threadIdx.x + 4503599627370000.0; lo[i] = i; hi[i] = threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 * (2 * i + 1) + blockIdx.x + 4503599627300000.0;
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < INNER; ++r) {
            // a fresh multiplier per repetition, derived from the accumulators (as a quotient digit would be), so that
            // ptxas cannot reuse products across repetitions: an integer < 2^52 as a double
            // (only the even accumulators are updated below: an odd one would make bb loop-invariant and let ptxas hoist
            // 3/8 of the products out of the loop -- the first version of this tool did, and overstated the rate by 1.6x)
            const uint64_t qb = lo[(r % (ACC / 2)) * 2];
            const double bb = __hiloint2double((int)((uint32_t)(qb >> 32) & 0xfffffu) | 0x43200000, (int)(uint32_t)qb) - 0x1p51 + b[r & 3] * 0.0;
#pragma unroll
            for (int i = 0; i < ACC; i += 2) {
                const double h0 = __fma_rz(a[i], bb, C1), h1 = __fma_rz(a[i + 1], bb, C1);
                const double l0 = __fma_rz(a[i], bb, C2 - h0), l1 = __fma_rz(a[i + 1], bb, C2 - h1);
                // column sums: each accumulator takes two values per step (3-input IADD3 / IADD3.X)
                hi[i] += (uint64_t)__double_as_longlong(h0) + (uint64_t)__double_as_longlong(l1);
                lo[i] += (uint64_t)__double_as_longlong(l0) + (uint64_t)__double_as_longlong(h1);
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= lo[i] ^ hi[i];
    if (s == 0x1234567u) out[0] = s;
}

// DFMA and IMAD.WIDE side by side: per repetition ACC DFMA (pure chains) and NW IMAD.WIDE (pure chains)
template <int NW>
__global__ void k_mixed(uint64_t* out, double a0, double b0, int trips) {
    double acc[ACC], a[ACC], b[4];
    uint32_t wl[ACC], wh[ACC], wa[ACC], wb[4];
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        acc[i] = i + threadIdx.x; a[i] = a0 * (i + 3) + threadIdx.x;
        wl[i] = i + threadIdx.x; wh[i] = 2 * i; wa[i] = 3u * (i + 3) + threadIdx.x;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = b0 * (2 * i + 1) + blockIdx.x; wb[i] = (5u ^ blockIdx.x) * (2 * i + 1) + 7; }
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < INNER; ++r) {
#pragma unroll
            for (int i = 0; i < ACC; ++i) {
                acc[i] = __fma_rz(a[i], b[r & 3], acc[i]);
                if (i < NW)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(wl[i]), "+r"(wh[i]) : "r"(wa[i]), "r"(wb[r & 3]));
            }
        }
    }
    double s = 0;
    uint64_t u = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) { s += acc[i]; u ^= wl[i] ^ ((uint64_t)wh[i] << 32); }
    if (s == 1.2345 || u == 0x1234567u) out[0] = (uint64_t)s + u;
}

template <typename K>
static double run(K kern, int blocks, int threads, int trips, uint64_t* d_out) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) kern<<<blocks, threads>>>(d_out, 3.0, 5.0, trips);
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<blocks, threads>>>(d_out, 3.0, 5.0, trips);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return (double)blocks * threads * (double)trips * INNER * ACC / (best * 1e-3);   // "ACC units per repetition" per second
}

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint64_t* d_out; CK(cudaMalloc(&d_out, 8));
    int clock_khz = 0; cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int trips = 1000;
    double best_dfma = 0, best_prod = 0, best_mix4 = 0, best_mix8 = 0;
    for (int wps = 4; wps <= 32; wps *= 2) {
        const int threads = 128, blocks = sms * (wps * 32 / threads);
        const double d = run(k_dfma_pure, blocks, threads, trips, d_out);
        const double p = run(k_dfma_product, blocks, threads, trips, d_out);
        const double m4 = run(k_mixed<4>, blocks, threads, trips, d_out);
        const double m8 = run(k_mixed<8>, blocks, threads, trips, d_out);
        fprintf(stderr, "warps/SM=%2d  DFMA %.3f T/s   limb products (52x52) %.3f T/s   mixed: DFMA %.3f T/s + IMAD.WIDE %.3f T/s | DFMA %.3f + IMAD.WIDE %.3f\n",
                wps, d / 1e12, p / 1e12, m4 / 1e12, m4 / 2e12, m8 / 1e12, m8 / 1e12);
        if (d > best_dfma) best_dfma = d;
        if (p > best_prod) best_prod = p;
        if (m4 > best_mix4) best_mix4 = m4;
        if (m8 > best_mix8) best_mix8 = m8;
    }
    char buf[1024];
    snprintf(buf, sizeof buf,
             "{\"gpu\": \"%s\", \"sms\": %d, \"max_sm_khz\": %d, \"dfma_tops\": %.4f, \"dfma_per_sm_per_clk\": %.2f, \"limb_products_52_tops\": %.4f, "
             "\"limb_product_pbit2_per_s\": %.3f, \"imad_wide_pbit2_per_s_at_9p25T\": %.3f, "
             "\"mixed_dfma_with_half_imadwide_tops\": %.4f, \"mixed_dfma_with_equal_imadwide_tops\": %.4f}",
             prop.name, sms, clock_khz, best_dfma / 1e12, best_dfma / sms / (clock_khz * 1e3), best_prod / 1e12,
             best_prod * 2704.0 / 1e15, 9.25e12 * 1024.0 / 1e15, best_mix4 / 1e12, best_mix8 / 1e12);
    printf("%s\n", buf);
    if (argc > 1) { FILE* f = fopen(argv[1], "w"); if (f) { fprintf(f, "%s\n", buf); fclose(f); } }
    return 0;
}
