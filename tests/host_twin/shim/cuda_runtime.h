// Host stand-in for <cuda_runtime.h>, used ONLY by tests/host_twin/twin.cpp: it lets g++ compile the device
// arithmetic headers (pyhype_b200/csrc/pyh_math.cuh, pyh_fastdiv.cuh) as plain C++ so that the formulas the
// stage kernel executes can be compared with the oracle on the CPU (tests/test_host_twin.py).  Test
// infrastructure: nothing under pyhype_b200/ includes or links it.
#pragma once
// every standard header first: the CUDA spellings defined below (__noinline__ ...) also occur inside libstdc++
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict

using std::fabs;
using std::floor;
using std::fma;
using std::ldexp;
using std::sqrt;

static inline int __double2hiint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)b; }
static inline double __hiloint2double(int hi, int lo) {
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
static inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
// CUDA's global min / max overloads (the ones the kernels use; an unsigned-only `max` would silently mangle max(j, -1))
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }

// Stand-ins for the MUFU.RCP64H / MUFU.RSQ64H seeds: they see only the high word of the operand and
// return a high word (low word zero).  The tables of the hardware unit are not reproduced -- any seed
// with ~2^-19 relative accuracy makes the refinement sequences of pyh_fastdiv.cuh converge to the same
// correctly rounded quotient / reciprocal / root, which is the property the twin tests.
namespace pyh_host_twin {
static inline int rcp64h(int hi) { return __double2hiint(1.0 / __hiloint2double(hi, 0)); }
static inline int rsq64h(int hi) { return __double2hiint(1.0 / std::sqrt(__hiloint2double(hi, 0))); }
}  // namespace pyh_host_twin

// ---- CTA emulation for tests/host_twin/kernel_twin.cpp: the stage kernel itself runs on the host, one OS thread per
// CUDA thread of a thread block, __syncthreads() == a std::barrier over the block; thread blocks run one after another.
#define __launch_bounds__(...)
// static __shared__ arrays become function-local statics: one copy shared by all threads, which is what they are inside a
// thread block (blocks run one after another).  The stage kernel's dynamic `extern __shared__` array is declared through
// PYH_HOST_TWIN in pyh_stage_march.cuh (an `extern static` does not exist).
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline thread_local dim3 threadIdx;
inline dim3 blockIdx, blockDim, gridDim;

namespace pyh_host_twin {
inline std::barrier<>* cta_barrier = nullptr;
// warp shuffles: the 32 threads of a warp meet at their own barrier around an exchange buffer
inline thread_local std::barrier<>* warp_barrier = nullptr;
inline thread_local unsigned long long* warp_buf = nullptr;   // 32 slots of this thread's warp

// run `body()` for every thread of every block of the grid; with_barrier spawns real threads (needed as soon as the
// kernel calls __syncthreads), otherwise the threads of a block run one after another on the calling thread
inline void launch(dim3 grid, unsigned nthreads, bool with_barrier, const std::function<void()>& body) {
    gridDim = grid;
    blockDim = dim3(nthreads);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = dim3(bx, by, bz);
                if (!with_barrier) {
                    for (unsigned t = 0; t < nthreads; ++t) { threadIdx = dim3(t); body(); }
                    continue;
                }
                std::barrier<> bar((std::ptrdiff_t)nthreads);
                cta_barrier = &bar;
                const unsigned nwarps = (nthreads + 31) / 32;
                std::vector<std::unique_ptr<std::barrier<>>> wbar;
                for (unsigned w = 0; w < nwarps; ++w)
                    wbar.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(32u, nthreads - 32 * w)));
                std::vector<unsigned long long> wbuf(32 * (size_t)nwarps, 0ull);
                std::vector<std::thread> th;
                th.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t)
                    th.emplace_back([t, &body, &bar, &wbar, &wbuf] {
                        threadIdx = dim3(t);
                        warp_barrier = wbar[t / 32].get();
                        warp_buf = wbuf.data() + 32 * (size_t)(t / 32);
                        body();
                        bar.arrive_and_drop();   // a thread that returns early must not hold the others
                        warp_barrier->arrive_and_drop();
                    });
                for (auto& x : th) x.join();
                cta_barrier = nullptr;
            }
}
}  // namespace pyh_host_twin

static inline void __syncthreads() { pyh_host_twin::cta_barrier->arrive_and_wait(); }
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) {}
static inline double __longlong_as_double(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
// full-warp butterfly exchange (every lane of the warp must call it, as in k_dt)
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "shuffle of up to 64 bits");
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    pyh_host_twin::warp_buf[lane] = bits;
    pyh_host_twin::warp_barrier->arrive_and_wait();
    const unsigned long long other = pyh_host_twin::warp_buf[lane ^ (unsigned)lane_mask];
    pyh_host_twin::warp_barrier->arrive_and_wait();
    T r;
    std::memcpy(&r, &other, sizeof(T));
    return r;
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = v; return o; }
