// Host emulation of the handful of CUDA device facilities paillier_b200/csrc/mont.cuh and mont52.cuh use, so that the
// multiplier SOURCES the kernels are built from can run in the CPU test suite (tests/cpp/mont_host_test.cpp): a warp is 32
// real threads, one per lane; every warp collective (__shfl_*_sync, __ballot_sync, __syncwarp) is a barrier of the 32
// threads, "shared memory" is an ordinary array they share.  Because the lanes are real threads, a missing __syncwarp()
// between a lane's shared-memory write and another lane's read is a genuine data race here -- ThreadSanitizer reports it
// (tools/host_sanitize.sh), which is the check compute-sanitizer's racecheck would make on the device.
// TEST INFRASTRUCTURE ONLY: nothing in the product includes this file.
#pragma once
#include <barrier>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define PGPU_HOST_EMULATION 1
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))

struct uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }

struct HostDim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local HostDim3 threadIdx, blockIdx;          // one emulated block = one warp: blockIdx.x = 0
inline const HostDim3 blockDim{32, 1, 1};

namespace hostwarp {
struct Warp {
    std::barrier<> bar{32};
    uint64_t slot[32] = {};
};
inline thread_local Warp* warp = nullptr;
inline thread_local int lane = 0;
inline uint4* shared_mem = nullptr;                        // the block's dynamic shared memory (vm_run.cuh: vm_smem)

// every lane deposits `raw`, all meet, every lane reads the slot it wants, all meet again (so that the next collective
// cannot overwrite a slot somebody has not read yet)
inline uint64_t exchange(uint64_t raw, int src) {
    warp->slot[lane] = raw;
    warp->bar.arrive_and_wait();
    const uint64_t got = warp->slot[src];
    warp->bar.arrive_and_wait();
    return got;
}
template <class T> inline uint64_t to_raw(T v) { static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits"); uint64_t r = 0; std::memcpy(&r, &v, sizeof(T)); return r; }
template <class T> inline T from_raw(uint64_t r) { T v; std::memcpy(&v, &r, sizeof(T)); return v; }

// run body(lane) on the 32 lanes of one emulated warp
inline void run_warp(const std::function<void(int)>& body) {
    Warp w;
    std::vector<std::thread> th;
    for (int l = 0; l < 32; ++l)
        th.emplace_back([&w, &body, l] { warp = &w; lane = l; threadIdx.x = (unsigned)l; body(l); });
    for (auto& t : th) t.join();
}
}  // namespace hostwarp

template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int base = hostwarp::lane & ~(width - 1);
    return hostwarp::from_raw<T>(hostwarp::exchange(hostwarp::to_raw(v), base + (src & (width - 1))));
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int base = hostwarp::lane & ~(width - 1);
    int src = hostwarp::lane - (int)delta;
    if (src < base) src = hostwarp::lane;                       // out of the segment: the lane keeps its own value
    return hostwarp::from_raw<T>(hostwarp::exchange(hostwarp::to_raw(v), src));
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int base = hostwarp::lane & ~(width - 1);
    int src = hostwarp::lane + (int)delta;
    if (src >= base + width) src = hostwarp::lane;
    return hostwarp::from_raw<T>(hostwarp::exchange(hostwarp::to_raw(v), src));
}
inline unsigned __ballot_sync(unsigned, bool pred) {
    hostwarp::warp->slot[hostwarp::lane] = pred ? 1u : 0u;
    hostwarp::warp->bar.arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)hostwarp::warp->slot[l] << l;
    hostwarp::warp->bar.arrive_and_wait();
    return m;
}
inline void __syncwarp(unsigned = 0xffffffffu) { hostwarp::warp->bar.arrive_and_wait(); }

inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t shift) {      // high word of (hi:lo) << (shift & 31)
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((v << (shift & 31)) >> 32);
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {      // low word of (hi:lo) >> (shift & 31)
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (shift & 31));
}
template <class T> inline T __ldg(const T* p) { return *p; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
// fused multiply-add rounded toward zero / subtraction rounded to nearest (compile with -frounding-math)
inline double __fma_rz(double a, double b, double c) {
    const int old = std::fegetround();
    std::fesetround(FE_TOWARDZERO);
    volatile double va = a, vb = b, vc = c;
    volatile double r = std::fma(va, vb, vc);
    std::fesetround(old);
    return r;
}
inline double __dsub_rn(double a, double b) {
    volatile double va = a, vb = b;
    volatile double r = va - vb;
    return r;
}
