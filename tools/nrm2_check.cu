// nrm2_x87_dd2 (double-double emulation of every x87 rounding, near-tie inputs flagged) vs the exact integer emulation of
// OpenBLAS dnrm2's x87 sequence, on random 4-vectors with mixed magnitudes.  Build & run on a GPU box:
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I include -o tools/nrm2_check tools/nrm2_check.cu && tools/nrm2_check
#include <cstdio>
#include <cuda_runtime.h>
#include "../pyhype_b200/csrc/pyh_math.cuh"
using namespace pyh;

__device__ unsigned long long rng(unsigned long long& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

// mode 0: uniform mantissas, exponents spread over +-spread around 0; mode 1: near-tie hunting (perfect squares + tiny)
__global__ void k(unsigned long long seed, int iters, int spread, int mode, unsigned long long* cnt) {
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long fast = 0, bad = 0;
    for (int it = 0; it < iters; ++it) {
        double x[4];
        for (int k2 = 0; k2 < 4; ++k2) {
            unsigned long long r = rng(s);
            double m = 1.0 + (double)(r >> 12) * 2.220446049250313e-16;
            int ex = (int)(rng(s) % (unsigned)(2 * spread + 1)) - spread;
            double v = ldexp(m, ex);
            if (mode == 1) {
                // small integers: sums of squares are often perfect squares or exactly representable midpoints
                v = (double)(rng(s) % 4096u) * ldexp(1.0, ex / 8);
            }
            if ((r & 7) == 0 && mode != 1) v = 0.0;
            if (r & 8) v = -v;
            x[k2] = v;
        }
        // the 2-wide branch-free evaluation the HLLL kernel uses: this vector next to a permuted / scaled copy
        double y[4] = {x[3] * 0.75, x[0], -x[1], x[2] * 3.0}, f2[2];
        bool ok2[2];
        nrm2_x87_dd2(x, y, f2, ok2);
        if (ok2[0]) { ++fast; if (f2[0] != nrm2_x87(x)) ++bad; }
        if (ok2[1]) { ++fast; if (f2[1] != nrm2_x87(y)) ++bad; }
    }
    atomicAdd(&cnt[0], fast);
    atomicAdd(&cnt[1], bad);
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 16);
    const int spreads[] = {0, 3, 20, 60, 120, 300};
    for (int mode = 0; mode < 2; ++mode)
        for (int sp : spreads) {
            cudaMemset(d, 0, 16);
            const int grid = 148 * 8, thr = 256, iters = 2000;
            k<<<grid, thr>>>(12345 + sp + 1000 * mode, iters, sp, mode, d);
            unsigned long long h[2];
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            double tot = 2.0 * (double)grid * thr * iters;
            printf("mode %d exponent spread +-%3d: %.3e vectors, fast path %.3f %%, mismatches among fast results: %llu\n",
                   mode, sp, tot, 100.0 * h[0] / tot, h[1]);
        }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
