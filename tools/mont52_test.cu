// Stand-alone check and rate measurement of the FP64-pipe Montgomery multiplier (paillier_b200/csrc/mont52.cuh)
// against host big-integer arithmetic (bn_host.hpp) and against the IMAD.WIDE multiplier (mont.cuh).
//
//   parity : x = a*b; 40x (x = x*x); x = x + a; x = x - b; x = x*a; x = x*1 -> record, all Montgomery products,
//            compared with the same sequence in BigU arithmetic mod n (R = 2^(52*S52)); records at the edges
//            (0, 1, n-1, 2^(32*S)-1) included.
//   rate   : K dependent multiplications per group over a full persistent grid, both multipliers.
//
// Build: make -C tools mont52_test        Run: tools/mont52_test [json-out]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../paillier_b200/csrc/bn_host.hpp"
#include "../paillier_b200/csrc/mont52.cuh"

using namespace pgpu;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

template <int TPI, int L, int S32>
__global__ void __launch_bounds__(128) k_parity(const uint32_t* mod, uint32_t np0, const uint32_t* one, const uint32_t* a, const uint32_t* b,
                                                uint32_t* out, int n_items, int nsq) {
    using M = Mont52<TPI, L, S32>;
    M m;
    m.init(mod, np0);
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) / TPI;
    const int item = group < n_items ? group : 0;
    double x[L], y[L], z[L], o[L];
    m.load_rec(y, a + (size_t)item * S32, S32);
    m.load_rec(z, b + (size_t)item * S32, S32);
    m.load_rec(o, one, S32);
    m.mul(x, y, z);
    for (int i = 0; i < nsq; ++i) m.sqr(x, x);
    m.mul(y, y, o);        // a * R^-1: a value < 2n whatever the record held
    m.mul(z, z, o);
    m.add(x, x, y);
    m.sub(x, x, z);
    m.mul(x, x, y);
    m.mul(x, x, o);
    if (group < n_items) m.store_rec(out + (size_t)item * S32, x, S32);
}

template <int TPI, int L, int S32>
__global__ void __launch_bounds__(128) k_rate52(const uint32_t* mod, uint32_t np0, const uint32_t* a, uint32_t* out, int nmul) {
    using M = Mont52<TPI, L, S32>;
    M m;
    m.init(mod, np0);
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) / TPI;
    double x[L], y[L];
    m.load_rec(y, a + (size_t)(group & 63) * S32, S32);
#pragma unroll
    for (int k = 0; k < L; ++k) x[k] = y[k];
#pragma unroll 1
    for (int i = 0; i < nmul; ++i) m.mul(x, x, y);
    m.store_rec(out + (size_t)group * S32, x, S32);
}

template <int TPI, int L>
__global__ void __launch_bounds__(128) k_rate32(const uint32_t* mod, uint32_t np0, const uint32_t* a, uint32_t* out, int nmul) {
    using M = Mont<TPI, L, false>;
    constexpr int S = TPI * L;
    M m;
    m.init(mod, np0);
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) / TPI;
    const int lt = (threadIdx.x & 31) & (TPI - 1);
    uint32_t x[L], y[L];
#pragma unroll
    for (int k = 0; k < L; ++k) { y[k] = a[(size_t)(group & 63) * S + lt * L + k]; x[k] = y[k]; }
#pragma unroll 1
    for (int i = 0; i < nmul; ++i) m.mul(x, x, y);
#pragma unroll
    for (int k = 0; k < L; ++k) out[(size_t)group * S + lt * L + k] = x[k];
}

static BigU random_big(std::mt19937_64& g, int limbs) {
    std::vector<uint32_t> v(limbs);
    for (auto& x : v) x = (uint32_t)g();
    return BigU::from_limbs(v.data(), v.size());
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static std::string g_json;

template <int TPI, int L, int S32, int TPI32, int L32>
static bool run_shape(int sms, int mod_bits) {
    using M = Mont52<TPI, L, S32>;
    std::mt19937_64 g(20260101u + S32);
    // ---- modulus: odd, exactly mod_bits bits
    BigU N = random_big(g, S32);
    N = N % BigU::pow2(mod_bits);
    N = N + BigU::pow2(mod_bits - 1);
    if (!N.is_odd()) N = N + BigU(1);
    N = N % BigU::pow2(mod_bits);
    if (N.bitlen() != (size_t)mod_bits) N = N + BigU::pow2(mod_bits - 1);
    const BigU R = BigU::pow2(M::RBITS);
    BigU Rinv;
    if (!BigU::modinv(R % N, N, Rinv)) { fprintf(stderr, "R not invertible\n"); return false; }
    const uint32_t np0 = mont_np0(N.v[0]);
    // ---- inputs
    const int n_items = 600;
    std::vector<BigU> A(n_items), B(n_items);
    const BigU full = BigU::pow2(32 * S32) - BigU(1);
    for (int i = 0; i < n_items; ++i) { A[i] = random_big(g, S32); B[i] = random_big(g, S32); }
    A[0] = BigU(0); B[1] = BigU(0); A[2] = BigU(1); B[2] = BigU(1); A[3] = N - BigU(1); B[3] = N - BigU(1);
    A[4] = full; B[4] = full; A[5] = N; B[5] = BigU(1); A[6] = N + BigU(1); B[6] = N - BigU(1); A[7] = full; B[7] = BigU(0);
    for (int i = 8; i < 300; ++i) { A[i] = A[i] % N; B[i] = B[i] % N; }
    std::vector<uint32_t> ha((size_t)n_items * S32), hb((size_t)n_items * S32);
    for (int i = 0; i < n_items; ++i) {
        auto la = A[i].limbs(S32), lb = B[i].limbs(S32);
        std::copy(la.begin(), la.end(), ha.begin() + (size_t)i * S32);
        std::copy(lb.begin(), lb.end(), hb.begin() + (size_t)i * S32);
    }
    uint32_t *d_mod, *d_one, *d_a, *d_b, *d_out;
    CK(cudaMalloc(&d_mod, S32 * 4)); CK(cudaMalloc(&d_one, S32 * 4));
    CK(cudaMalloc(&d_a, ha.size() * 4)); CK(cudaMalloc(&d_b, hb.size() * 4));
    auto lm = N.limbs(S32), lo = BigU(1).limbs(S32);
    CK(cudaMemcpy(d_mod, lm.data(), S32 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_one, lo.data(), S32 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_a, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    const int gpb = 128 / TPI;
    const int grid_groups = sms * 4 * gpb;
    CK(cudaMalloc(&d_out, (size_t)std::max(grid_groups * 2, n_items) * S32 * 4));
    CK(cudaMemset(d_out, 0xee, (size_t)n_items * S32 * 4));
    const int nsq = 40;
    k_parity<TPI, L, S32><<<(n_items + gpb - 1) / gpb, 128>>>(d_mod, np0, d_one, d_a, d_b, d_out, n_items, nsq);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> hout((size_t)n_items * S32);
    CK(cudaMemcpy(hout.data(), d_out, hout.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    auto mm = [&](const BigU& x, const BigU& y) { return (((x * y) % N) * Rinv) % N; };
    for (int i = 0; i < n_items; ++i) {
        BigU x = mm(A[i], B[i]);
        for (int s = 0; s < nsq; ++s) x = mm(x, x);
        const BigU y = mm(A[i], BigU(1)), z = mm(B[i], BigU(1));
        x = (x + y) % N;
        x = (x + N - z) % N;
        x = mm(x, y);
        x = mm(x, BigU(1));
        const BigU got = BigU::from_limbs(hout.data() + (size_t)i * S32, S32);
        if (!(got == x)) { if (bad < 5) fprintf(stderr, "shape %dx%d item %d differs\n got %s\n exp %s\n", TPI, L, i, got.hex().c_str(), x.hex().c_str()); ++bad; }
    }
    fprintf(stderr, "Mont52<%d,%d> S32=%d (%d-bit modulus): parity %s (%d of %d differ)\n", TPI, L, S32, mod_bits, bad ? "FAILED" : "ok", bad, n_items);
    // ---- rate
    int occ52 = 0, occ32 = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ52, k_rate52<TPI, L, S32>, 128, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, k_rate32<TPI32, L32>, 128, 0));
    const int nmul = 2000;
    const int b52 = sms * occ52, b32 = sms * occ32;
    const double ms52 = time_ms([&] { k_rate52<TPI, L, S32><<<b52, 128>>>(d_mod, np0, d_a, d_out, nmul); });
    const double ms32 = time_ms([&] { k_rate32<TPI32, L32><<<b32, 128>>>(d_mod, np0, d_a, d_out, nmul); });
    const double r52 = (double)b52 * gpb * nmul / (ms52 * 1e-3), r32 = (double)b32 * (128 / TPI32) * nmul / (ms32 * 1e-3);
    cudaFuncAttributes fa52, fa32;
    CK(cudaFuncGetAttributes(&fa52, k_rate52<TPI, L, S32>)); CK(cudaFuncGetAttributes(&fa32, k_rate32<TPI32, L32>));
    fprintf(stderr, "  rate: FP64 %dx%d %.1f M mul/s (%d regs, %d blocks/SM)   IMAD %dx%d %.1f M mul/s (%d regs, %d blocks/SM)   ratio %.3f\n",
            TPI, L, r52 / 1e6, fa52.numRegs, occ52, TPI32, L32, r32 / 1e6, fa32.numRegs, occ32, r52 / r32);
    // ---- both multipliers at once: `a` integer-pipe blocks and `b` FP64-pipe blocks per SM on two streams (do the IMAD.WIDE
    // pipe and the FP64 + ALU pipes overlap across warps?)
    if (getenv("MONT52_MIX")) {
        cudaStream_t s1, s2;
        CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
        uint32_t* d_out2; CK(cudaMalloc(&d_out2, (size_t)grid_groups * 4 * S32 * 4));
        const int combos[][2] = {{1, 1}, {1, 2}, {2, 1}, {1, 3}, {2, 2}};
        for (auto& cb : combos) {
            const int a = cb[0], b = cb[1];
            if (a > occ32 || b > occ52) continue;
            // multiplications per kernel in proportion to its stand-alone rate per block, so that both finish together
            const double per_block32 = r32 / (double)(sms * occ32), per_block52 = r52 / (double)(sms * occ52);
            const int n32 = 1500, n52 = (int)(n32 * (per_block52 / (128 / TPI)) / (per_block32 / (128 / TPI32)));
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            double best = 1e30;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, s1));
                CK(cudaStreamWaitEvent(s2, e0, 0));
                k_rate32<TPI32, L32><<<sms * a, 128, 0, s1>>>(d_mod, np0, d_a, d_out, n32);
                k_rate52<TPI, L, S32><<<sms * b, 128, 0, s2>>>(d_mod, np0, d_a, d_out2, n52);
                CK(cudaEventRecord(e1, s2));
                CK(cudaStreamWaitEvent(s1, e1, 0));
                CK(cudaEventRecord(e1, s1));
                CK(cudaEventSynchronize(e1));
                float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            const double muls = (double)sms * a * (128 / TPI32) * n32 + (double)sms * b * gpb * n52;
            fprintf(stderr, "  mix: %d integer + %d FP64 blocks/SM: %.1f M mul/s combined (%.3f x the better one alone)\n", a, b, muls / (best * 1e-3) / 1e6,
                    muls / (best * 1e-3) / std::max(r52, r32));
        }
        cudaFree(d_out2);
    }
    char buf[512];
    snprintf(buf, sizeof buf, "%s{\"mod_bits\": %d, \"fp64_shape\": \"%dx%d\", \"fp64_mmul_per_s\": %.2f, \"fp64_regs\": %d, \"fp64_blocks_per_sm\": %d, "
             "\"imad_shape\": \"%dx%d\", \"imad_mmul_per_s\": %.2f, \"imad_regs\": %d, \"ratio\": %.4f, \"parity_mismatches\": %d}",
             g_json.empty() ? "" : ", ", mod_bits, TPI, L, r52 / 1e6, fa52.numRegs, occ52, TPI32, L32, r32 / 1e6, fa32.numRegs, r52 / r32, bad);
    g_json += buf;
    cudaFree(d_mod); cudaFree(d_one); cudaFree(d_a); cudaFree(d_b); cudaFree(d_out);
    return bad == 0;
}

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    bool ok = true;
    const int only = argc > 2 ? atoi(argv[2]) : -1;     // run one shape only (for ncu)
    int idx = 0;
#define SHAPE(A, B, C, D, E, BITS) do { if (only < 0 || only == idx) ok &= run_shape<A, B, C, D, E>(sms, BITS); ++idx; } while (0)
    SHAPE(4, 5, 32, 4, 8, 1024);
    SHAPE(4, 10, 64, 4, 16, 2048);
    SHAPE(8, 5, 64, 4, 16, 2048);
    SHAPE(4, 15, 96, 4, 24, 3072);
    SHAPE(8, 8, 96, 8, 12, 3072);
    SHAPE(8, 10, 128, 4, 32, 4096);
    SHAPE(16, 5, 128, 4, 32, 4096);
    SHAPE(4, 20, 128, 4, 32, 4096);
    SHAPE(8, 15, 192, 8, 24, 6144);
    SHAPE(16, 8, 192, 8, 24, 6144);
#undef SHAPE
    std::string out = std::string("{\"gpu\": \"") + prop.name + "\", \"shapes\": [" + g_json + "], \"all_parity_ok\": " + (ok ? "true" : "false") + "}";
    printf("%s\n", out.c_str());
    if (argc > 1) { FILE* f = fopen(argv[1], "w"); if (f) { fprintf(f, "%s\n", out.c_str()); fclose(f); } }
    return ok ? 0 : 1;
}
