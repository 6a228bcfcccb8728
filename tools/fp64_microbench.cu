// FP64 pipe micro-benchmark for B200 (sm_100a): pins the FP64 roofline used in DESIGN.md.
// Measures issue rates of DFMA, unfused DMUL+DADD, IEEE division, IEEE sqrt, DMNMX and a double2 copy.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int OP> __global__ void __launch_bounds__(256) k(double* out, double a0, double b0, int iters){
    double x[8];
    #pragma unroll
    for(int i=0;i<8;i++) x[i]=a0+threadIdx.x*1e-3+i;
    double b=b0, c=1.0000001;
    for(int it=0; it<iters; ++it){
        #pragma unroll
        for(int i=0;i<8;i++){
            if(OP==0) x[i]=fma(x[i],c,b);
            else if(OP==1) x[i]=__dadd_rn(__dmul_rn(x[i],c),b);
            else if(OP==2) x[i]=__ddiv_rn(b,x[i])+2.0;
            else if(OP==3) x[i]=__dsqrt_rn(x[i])+3.0;
            else if(OP==4) x[i]=fmax(fmin(x[i],b),c)+1e-9;
            else if(OP==5) x[i]=__dmul_rn(x[i],c);
            else if(OP==6) x[i]=__dadd_rn(x[i],c);
        }
    }
    double s=0; 
    #pragma unroll
    for(int i=0;i<8;i++) s+=x[i];
    if(s==123.456) out[0]=s;
}
__global__ void copyk(const double2* __restrict__ a, double2* __restrict__ b, size_t n){
    size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x;
    for(;i<n;i+=st) b[i]=a[i];
}
template<int OP> int run(const char* name, double ops_per_iter_elem){
    double* d; CK(cudaMalloc(&d,8));
    int iters=4096; int grid=148*8, blk=256;
    k<OP><<<grid,blk>>>(d,1.5,2.5,16); CK(cudaDeviceSynchronize());
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best=1e30f;
    for(int r=0;r<5;r++){ cudaEventRecord(e0); k<OP><<<grid,blk>>>(d,1.5,2.5,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    double n=(double)grid*blk*iters*8.0*ops_per_iter_elem;
    printf("%-28s %8.3f ms  %10.3f Gop/s (thread-level ops)\n",name,best,n/best*1e-6);
    cudaFree(d); return 0;
}
int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n",p.name,p.major,p.minor,p.multiProcessorCount,p.clockRate);
    run<0>("DFMA",1); run<1>("DMUL+DADD (2 instr)",2); run<5>("DMUL",1); run<6>("DADD",1);
    run<2>("DDIV_rn (+1 DADD)",1); run<3>("DSQRT_rn (+1 DADD)",1); run<4>("DMNMX x2 (+1 DADD)",1);
    size_t n=(size_t)1<<27; double2 *a,*b; CK(cudaMalloc(&a,n*16)); CK(cudaMalloc(&b,n*16)); CK(cudaMemset(a,1,n*16));
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for(int g=1; g<=16; g*=2){ float best=1e30f; for(int r=0;r<5;r++){ cudaEventRecord(e0); copyk<<<148*g*2,512>>>(a,b,n); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;} printf("copy double2 2GiB+2GiB grid=148*%d: %.3f ms %.1f GB/s\n",g*2,best,2.0*n*16/best*1e-6);}    
    return 0;
}
