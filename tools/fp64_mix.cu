// FP64 issue model of B200 (sm_100a): does an integer / FP32-pipe instruction issue in the shadow of a DFMA (the FP64 pipe takes a
// warp instruction every 2 cycles per scheduler), or does every DFMA hold the scheduler's issue port for both cycles?
// The stage kernel's stream is 44 % FP64 and 56 % other instructions (profiles/r01s_summary.md): under model A
// (shadow issue) its floor is max(2 x 1152, 2612) = 2612 cycles per warp-row, under model B 2 x 1152 + 1460 = 3764.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_mix tools/fp64_mix.cu && tools/fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

// ILP independent DFMA chains; after each DFMA, NI independent integer instructions (LOP3 / IADD3 chains of their own)
template <int ILP, int NI, int KIND>
__global__ void k(double* out, double c, double b, int iters, int seed) {
    double x[ILP];
    unsigned a[ILP * (NI > 0 ? NI : 1)];
    float f[ILP * (NI > 0 ? NI : 1)];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = 1.0 + threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < ILP * (NI > 0 ? NI : 1); i++) { a[i] = threadIdx.x + i + seed; f[i] = (float)a[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(c), "d"(b));
#pragma unroll
            for (int n = 0; n < NI; n++) {
                if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i * NI + n]) : "r"(seed), "r"(it));      // ALU pipe
                else if (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i * NI + n]) : "f"(1.0001f), "f"(0.5f));   // FMA pipe
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i * NI + n]) : "r"(seed), "r"(it));                      // IMAD (FMA pipe)
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
#pragma unroll
    for (int i = 0; i < ILP * (NI > 0 ? NI : 1); i++) s += (double)a[i] + (double)f[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP, int NI, int KIND>
void run(int warps_per_sm, const char* what) {
    double* d;
    cudaMalloc(&d, 8);
    int threads = 128, ctas_per_sm = warps_per_sm / 4, grid = 148 * ctas_per_sm, iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<ILP, NI, KIND><<<grid, threads>>>(d, 1.0000001, 1e-9, 100, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<ILP, NI, KIND><<<grid, threads>>>(d, 1.0000001, 1e-9, iters, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per scheduler: warps_per_sm / 4 warps, each issuing iters * ILP DFMA (+ NI others each)
    double dfma_per_sched = (double)(warps_per_sm / 4) * iters * ILP;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-10s warps/SM=%2d ILP=%d others/DFMA=%d: %8.3f ms  %.2f cycles per DFMA per scheduler (%.2f per instruction)\n", what, warps_per_sm, ILP, NI,
           ms, cyc / dfma_per_sched, cyc / (dfma_per_sched * (1 + NI)));
    cudaFree(d);
}

int main() {
    printf("-- DFMA latency / throughput vs ILP and warps per scheduler\n");
    run<1, 0, 0>(4, "dfma");  run<2, 0, 0>(4, "dfma");  run<4, 0, 0>(4, "dfma");  run<8, 0, 0>(4, "dfma");
    run<1, 0, 0>(16, "dfma"); run<2, 0, 0>(16, "dfma"); run<4, 0, 0>(16, "dfma"); run<8, 0, 0>(16, "dfma");
    printf("-- 16 warps/SM (4 per scheduler, like the stage kernel), ILP 4: DFMA + n other instructions each\n");
    run<4, 1, 0>(16, "lop3");  run<4, 2, 0>(16, "lop3");  run<4, 3, 0>(16, "lop3");
    run<4, 1, 1>(16, "ffma");  run<4, 2, 1>(16, "ffma");  run<4, 3, 1>(16, "ffma");
    run<4, 1, 2>(16, "imad");  run<4, 2, 2>(16, "imad");  run<4, 3, 2>(16, "imad");
    printf("-- same with ILP 2 (dependent chains dominate)\n");
    run<2, 0, 0>(16, "dfma");  run<2, 1, 0>(16, "lop3");  run<2, 2, 0>(16, "lop3");
    printf("-- 32 warps/SM\n");
    run<2, 1, 0>(32, "lop3");  run<4, 1, 0>(32, "lop3");  run<4, 2, 0>(32, "lop3");
    return 0;
}
