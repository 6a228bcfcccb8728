// FP64 dependent-chain throughput vs ILP and warps/SM on B200: how much ILP x TLP saturates the FP64 pipe?
#include <cstdio>
#include <cuda_runtime.h>
template<int ILP> __global__ void k(double* out, double c, double b, int iters){
    double x[ILP];
    #pragma unroll
    for(int i=0;i<ILP;i++) x[i]=1.0+threadIdx.x*1e-3+i;
    for(int it=0; it<iters; ++it){
        #pragma unroll
        for(int i=0;i<ILP;i++) x[i]=fma(x[i],c,b);
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<ILP;i++) s+=x[i];
    if(s==123.456) out[0]=s;
}
template<int ILP> void run(int warps_per_sm){
    double* d; cudaMalloc(&d,8);
    int threads = 128; int ctas_per_sm = warps_per_sm/4; int grid=148*ctas_per_sm; int iters=20000;
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<grid,threads>>>(d,1.0000001,1e-9,100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<ILP><<<grid,threads>>>(d,1.0000001,1e-9,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double ops=(double)grid*threads*iters*ILP;
    // cycles per dependent op per warp: time * clock / iters
    printf("warps/SM=%2d ILP=%d: %7.2f Gop/s (%.1f%% of 18270)  -> %.2f cycles per chain step\n", warps_per_sm, ILP, ops/ms*1e-6, ops/ms*1e-6/18270*100, ms*1e-3*1.92e9/iters);
    cudaFree(d);
}
int main(){
    for(int w : {4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
