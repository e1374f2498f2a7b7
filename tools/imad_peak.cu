// Integer-multiply roofline microbenchmark for B200 (sm_100a).
//
// Measures the dependency-free issue rate of the instruction the Montgomery
// kernels are made of (IMAD.WIDE.U32: 32x32+64 -> 64) and, for comparison, of
// the same instruction in carry chains (IMAD.WIDE.U32.X, as emitted for
// mad.lo.cc/madc.hi.cc pairs) and of the plain 32-bit IMAD.  One IMAD.WIDE is
// one MAC32 in SURVEY.md section 8(d)'s accounting; the measured rate is the
// denominator of bench.py's roofline.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak imad_peak.cu
// Run:   ./imad_peak [json-out]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ACC = 8;       // independent 64-bit accumulators per thread
constexpr int INNER = 64;    // unrolled repetitions per loop trip

// Every repetition derives a fresh multiplier q from the accumulators (as the
// CIOS quotient digit does), so ptxas cannot hoist or strength-reduce the
// products; per repetition: 1 IMAD + ACC multiply-accumulates.
__global__ void k_wide_indep(uint64_t* out, uint32_t a0, uint32_t b0, int trips) {
    uint32_t lo[ACC], hi[ACC], a[ACC];
    const uint32_t np = b0 ^ blockIdx.x;
#pragma unroll
    for (int i = 0; i < ACC; ++i) { lo[i] = i + threadIdx.x; hi[i] = 2 * i; a[i] = a0 * (i + 3) + threadIdx.x; }
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < INNER; ++r) {
            const uint32_t q = lo[r % ACC] * np;
#pragma unroll
            for (int i = 0; i < ACC; ++i)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(q));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= lo[i] ^ ((uint64_t)hi[i] << 32);
    if (s == 0x1234567u) out[0] = s;
}

// the same multiply-accumulates with loop-invariant multipliers: nothing but IMAD.WIDE in the loop body
__global__ void k_wide_pure(uint64_t* out, uint32_t a0, uint32_t b0, int trips) {
    uint32_t lo[ACC], hi[ACC], a[ACC], b[4];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { lo[i] = i + threadIdx.x; hi[i] = 2 * i; a[i] = a0 * (i + 3) + threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (b0 ^ blockIdx.x) * (2 * i + 1) + 7;
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < INNER; ++r) {
#pragma unroll
            for (int i = 0; i < ACC; ++i)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(b[r & 3]));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= lo[i] ^ ((uint64_t)hi[i] << 32);
    if (s == 0x1234567u) out[0] = s;
}

// two carry chains of ACC/2 column pairs each, like one CIOS half-step
__global__ void k_wide_chain(uint64_t* out, uint32_t a0, uint32_t b0, int trips) {
    uint32_t lo[ACC], hi[ACC], a[ACC];
    const uint32_t np = b0 ^ blockIdx.x;
    uint32_t c0 = 0, c1 = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) { lo[i] = i + threadIdx.x; hi[i] = 2 * i; a[i] = a0 * (i + 3) + threadIdx.x; }
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < INNER; ++r) {
            const uint32_t q = lo[r % ACC] * np;
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a[0]), "r"(q));
#pragma unroll
            for (int i = 1; i < ACC / 2; ++i)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(q));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(c0));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[ACC / 2]), "+r"(hi[ACC / 2]) : "r"(a[ACC / 2]), "r"(q));
#pragma unroll
            for (int i = ACC / 2 + 1; i < ACC; ++i)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(q));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(c1));
        }
    }
    uint64_t s = c0 + c1;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= lo[i] ^ ((uint64_t)hi[i] << 32);
    if (s == 0x1234567u) out[0] = s;
}

__global__ void k_imad32(uint64_t* out, uint32_t a0, uint32_t b0, int trips) {
    uint32_t acc[ACC], a[ACC];
    const uint32_t np = b0 ^ blockIdx.x;
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i] = i + threadIdx.x; a[i] = a0 * (i + 3) + threadIdx.x; }
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < INNER; ++r) {
            const uint32_t q = acc[r % ACC] * np;
#pragma unroll
            for (int i = 0; i < ACC; ++i)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a[i]), "r"(q));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= acc[i];
    if (s == 0x1234567u) out[0] = s;
}

template <typename K>
static double run(K kern, int blocks, int threads, int trips, uint64_t* d_out) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) kern<<<blocks, threads>>>(d_out, 3, 5, trips);
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<blocks, threads>>>(d_out, 3, 5, trips);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    double ops = (double)blocks * threads * (double)trips * INNER * ACC;   // the extra IMAD per repetition is not counted
    return ops / (best * 1e-3);
}

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    uint64_t* d_out; CK(cudaMalloc(&d_out, 8));
    int clock_khz = 0; cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int trips = 2000;
    double best_wide = 0, best_chain = 0, best_32 = 0, best_pure = 0;
    int cfg_wide = 0;
    for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM
        int threads = 128, blocks = sms * (wps * 32 / threads);
        double w = run(k_wide_indep, blocks, threads, trips, d_out);
        double c = run(k_wide_chain, blocks, threads, trips, d_out);
        double pu = run(k_wide_pure, blocks, threads, trips, d_out);
        if (pu > best_pure) best_pure = pu;
        double s = run(k_imad32, blocks, threads, trips, d_out);
        fprintf(stderr, "warps/SM=%2d  IMAD.WIDE indep %.3f T/s  pure %.3f T/s  chain %.3f T/s  IMAD32 %.3f T/s\n", wps, w / 1e12, pu / 1e12, c / 1e12, s / 1e12);
        if (w > best_wide) { best_wide = w; cfg_wide = wps; }
        if (c > best_chain) best_chain = c;
        if (s > best_32) best_32 = s;
    }
    char buf[1024];
    snprintf(buf, sizeof buf,
             "{\"gpu\": \"%s\", \"sms\": %d, \"max_sm_khz\": %d, \"imad_wide_tmacs\": %.4f, \"imad_wide_with_q_tmacs\": %.4f, \"imad_wide_chain_tmacs\": %.4f, "
             "\"imad32_tops\": %.4f, \"best_warps_per_sm\": %d, \"per_sm_per_clk_at_max\": %.2f}",
             prop.name, sms, clock_khz, (best_pure > best_wide ? best_pure : best_wide) / 1e12, best_wide / 1e12, best_chain / 1e12, best_32 / 1e12, cfg_wide,
             (best_pure > best_wide ? best_pure : best_wide) / sms / (clock_khz * 1e3));
    printf("%s\n", buf);
    if (argc > 1) { FILE* f = fopen(argv[1], "w"); if (f) { fprintf(f, "%s\n", buf); fclose(f); } }
    return 0;
}
