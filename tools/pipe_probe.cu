// Which instructions issue beside a DFMA on sm_100a?  (r02: the FP64-pipe multiplier of mont52.cuh reaches only 58 % of
// the DFMA issue rate; ncu shows math-pipe-throttle and dispatch stalls on DFMA, DADD and IADD3 alike.)
//
// Every kernel runs F independent DFMA chains and A independent instructions of one other kind per repetition and the
// tool prints the cycles one scheduler (SM sub-partition) needs per repetition, next to 2*F (FP64 pipe alone) and to what
// the other instructions need alone.  If the two overlap the time is the maximum, if they share the dispatch port it is
// the sum.
//
// Build: make -C tools pipe_probe     Run: tools/pipe_probe [json-out]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

enum Kind { K_NONE, K_IADD, K_IADD3C, K_LOP3, K_IMAD, K_IMADW, K_FFMA, K_SHFL, K_DADD, K_IADDX_PAIR };

constexpr int REPS = 16;

template <int F, int A, int KIND>
__global__ void __launch_bounds__(128) k_mix(uint64_t* out, double c0, uint32_t seed, int trips) {
    double d[8], e[8];
    uint32_t u[16], v[16];
    float f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = c0 * (i + 1) + threadIdx.x; e[i] = c0 * 0.5 + i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) { u[i] = seed * (i + 3) + threadIdx.x; v[i] = seed + 7 * i; f[i] = (float)(seed + i); }
    const double cm = c0 * 1.0000001;
    const uint32_t w = seed | 1u;
    const float fm = 1.0000001f;
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int r = 0; r < REPS; ++r) {
#pragma unroll
            for (int i = 0; i < (F > A ? F : A); ++i) {
                if (i < F) d[i] = __fma_rz(d[i], cm, e[i]);
                if (i < A) {
                    if (KIND == K_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[i]));
                    if (KIND == K_IADD3C) {   // 64-bit three-input add: IADD3 with two carry-outs + IADD3.X
                        uint64_t x = ((uint64_t)u[i] << 32) | v[i];
                        x += (uint64_t)__double_as_longlong(d[i & 7]) + (uint64_t)__double_as_longlong(e[i & 7]);
                        u[i] = (uint32_t)(x >> 32); v[i] = (uint32_t)x;
                    }
                    if (KIND == K_IADDX_PAIR) {   // 64-bit two-input add: IADD3 + IADD3.X / IMAD.X
                        uint64_t x = ((uint64_t)u[i] << 32) | v[i];
                        x += (uint64_t)__double_as_longlong(e[i & 7]);
                        u[i] = (uint32_t)(x >> 32); v[i] = (uint32_t)x;
                    }
                    if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(w));
                    if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(w), "r"(v[i]));
                    if (KIND == K_IMADW) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(u[i]), "+r"(v[i]) : "r"(w), "r"(seed));
                    if (KIND == K_FFMA) f[i] = __fmaf_rn(f[i], fm, 0.5f);
                    if (KIND == K_SHFL) u[i] = __shfl_sync(0xffffffffu, u[i], (i + 1) & 31);
                    if (KIND == K_DADD) e[i & 7] = __dadd_rn(e[i & 7], cm);
                }
            }
        }
    }
    double s = 0; uint64_t x = 0; float g = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i] + e[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x ^= u[i] ^ ((uint64_t)v[i] << 32); g += f[i]; }
    if (s == 1.2345 || x == 0x1234567u || g == 1.5f) out[0] = (uint64_t)s + x;
}

static std::string g_json;
static int g_sms = 0, g_khz = 0;

template <int F, int A, int KIND>
static void run(const char* name, uint64_t* d_out, int warps_per_smsp) {
    const int trips = 400;
    const int threads = 128, blocks = g_sms * warps_per_smsp;   // 4 warps per block, one per scheduler
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) k_mix<F, A, KIND><<<blocks, threads>>>(d_out, 3.0, 12345u, trips);
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k_mix<F, A, KIND><<<blocks, threads>>>(d_out, 3.0, 12345u, trips);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    // cycles one scheduler spends per repetition of one warp
    const double cyc = best * 1e-3 * (g_khz * 1e3) / ((double)warps_per_smsp * trips * REPS);
    fprintf(stderr, "%-12s F=%d A=%2d warps/scheduler=%d : %7.2f cycles per repetition (DFMA alone %d)\n", name, F, A, warps_per_smsp, cyc, 2 * F);
    char buf[256];
    snprintf(buf, sizeof buf, "%s{\"kind\": \"%s\", \"dfma\": %d, \"other\": %d, \"warps_per_scheduler\": %d, \"cycles_per_rep\": %.3f}",
             g_json.empty() ? "" : ", ", name, F, A, warps_per_smsp, cyc);
    g_json += buf;
}

#define SWEEP(KIND, NAME)                                            \
    run<0, 8, KIND>(NAME, d_out, 4); run<0, 16, KIND>(NAME, d_out, 4);  \
    run<8, 4, KIND>(NAME, d_out, 4); run<8, 8, KIND>(NAME, d_out, 4); run<8, 16, KIND>(NAME, d_out, 4); \
    run<8, 8, KIND>(NAME, d_out, 8);

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    cudaDeviceGetAttribute(&g_khz, cudaDevAttrClockRate, 0);
    uint64_t* d_out; CK(cudaMalloc(&d_out, 8));
    run<8, 0, K_NONE>("dfma", d_out, 1); run<8, 0, K_NONE>("dfma", d_out, 2); run<8, 0, K_NONE>("dfma", d_out, 4); run<8, 0, K_NONE>("dfma", d_out, 8);
    SWEEP(K_IADD, "iadd")
    SWEEP(K_IADD3C, "iadd3_64x3")
    SWEEP(K_IADDX_PAIR, "iadd_64x2")
    SWEEP(K_LOP3, "lop3")
    SWEEP(K_IMAD, "imad")
    SWEEP(K_IMADW, "imad_wide")
    SWEEP(K_FFMA, "ffma")
    SWEEP(K_SHFL, "shfl")
    SWEEP(K_DADD, "dadd")
    std::string out = std::string("{\"gpu\": \"") + prop.name + "\", \"sm_khz\": " + std::to_string(g_khz) + ", \"runs\": [" + g_json + "]}";
    printf("%s\n", out.c_str());
    if (argc > 1) { FILE* f = fopen(argv[1], "w"); if (f) { fprintf(f, "%s\n", out.c_str()); fclose(f); } }
    return 0;
}
