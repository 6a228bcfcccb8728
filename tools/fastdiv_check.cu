#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "/root/repo/pyhype_b200/csrc/pyh_fastdiv.cuh"
using namespace pyh;

__device__ unsigned long long splitmix(unsigned long long& s){ s += 0x9E3779B97f4A7C15ull; unsigned long long z=s; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31); }
__device__ double rnd_double(unsigned long long& s, int mode){
    unsigned long long r = splitmix(s);
    if (mode == 0) { // random mantissa, exponent in [-40, 40]
        unsigned long long m = r & 0x000fffffffffffffull; int e = 1023 + (int)((r >> 52) % 81) - 40; unsigned long long sg = (r>>63)<<63;
        return __longlong_as_double((long long)(sg | ((unsigned long long)e << 52) | m));
    } else if (mode == 1) { // any bit pattern
        return __longlong_as_double((long long)r);
    } else { // near-1 values with few mantissa bits (hard rounding cases)
        unsigned long long m = (r & 0xfffffull) << 32 | (splitmix(s) & 0x7); return __longlong_as_double((long long)((1023ull<<52)|m));
    }
}
__global__ void test(unsigned long long seed, int iters, int mode, unsigned long long* bad, unsigned long long* fallback){
    unsigned long long s = seed + (blockIdx.x*(unsigned long long)blockDim.x + threadIdx.x) * 0x1234567ull;
    unsigned long long nbad=0, nfb=0;
    for(int it=0; it<iters; ++it){
        double a = rnd_double(s, mode), b = rnd_double(s, mode);
        if ((it & 15) == 0) a = 0.0;
        // division
        bool ok = true;
        Recip rb = recip_prepare(b, ok);
        double q = div_fast(a, rb, ok);
        double qe = a / b;
        if (ok) { if (!(q == qe) && !(q != q && qe != qe)) nbad++; } else nfb++;
        // plain reciprocal
        bool ok2 = true; double r1 = rcp_fast(b, ok2); double r1e = 1.0 / b;
        if (ok2) { if (!(r1 == r1e) && !(r1 != r1 && r1e != r1e)) nbad++; } else nfb++;
        // sqrt
        bool ok3 = true; double x = fabs(a); double sq = sqrt_fast(x, ok3); double sqe = sqrt(x);
        if (ok3) { if (!(sq == sqe)) nbad++; } else nfb++;
    }
    atomicAdd(bad, nbad); atomicAdd(fallback, nfb);
}
int main(){
    unsigned long long *bad, *fb; cudaMalloc(&bad,8); cudaMalloc(&fb,8);
    for(int mode=0; mode<3; ++mode){
        cudaMemset(bad,0,8); cudaMemset(fb,0,8);
        test<<<148*8,256>>>(12345+mode, 4000, mode, bad, fb);
        unsigned long long hb, hf; cudaMemcpy(&hb,bad,8,cudaMemcpyDeviceToHost); cudaMemcpy(&hf,fb,8,cudaMemcpyDeviceToHost);
        printf("mode %d: %llu mismatches, %llu fallbacks of %llu ops (%s)\n", mode, hb, hf, 3ull*148*8*256*4000, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
