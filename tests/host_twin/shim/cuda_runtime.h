// Host stand-in for <cuda_runtime.h>, used ONLY by tests/host_twin/twin.cpp: it lets g++ compile the device
// arithmetic headers (pyhype_b200/csrc/pyh_math.cuh, pyh_fastdiv.cuh) as plain C++ so that the formulas the
// stage kernel executes can be compared with the oracle on the CPU (tests/test_host_twin.py).  Test
// infrastructure: nothing under pyhype_b200/ includes or links it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

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
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }

// Stand-ins for the MUFU.RCP64H / MUFU.RSQ64H seeds: they see only the high word of the operand and
// return a high word (low word zero).  The tables of the hardware unit are not reproduced -- any seed
// with ~2^-19 relative accuracy makes the refinement sequences of pyh_fastdiv.cuh converge to the same
// correctly rounded quotient / reciprocal / root, which is the property the twin tests.
namespace pyh_host_twin {
static inline int rcp64h(int hi) { return __double2hiint(1.0 / __hiloint2double(hi, 0)); }
static inline int rsq64h(int hi) { return __double2hiint(1.0 / std::sqrt(__hiloint2double(hi, 0))); }
}  // namespace pyh_host_twin
