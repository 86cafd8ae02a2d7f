// FP64 pipe microbenchmark for B200 (sm_100a): DMMA.8x8x4 vs DFMA issue rates.
// Used once to fix the FP64 roofline denominator quoted in DESIGN.md (MEASURED_PEAKS.json has no FP64 figure).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void dmma_kernel(double* out, const double* in, int iters){
  double a=in[threadIdx.x], b=in[threadIdx.x+32];
  double c[ILP][2];
  #pragma unroll
  for(int j=0;j<ILP;j++){c[j][0]=0;c[j][1]=0;}
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<ILP;j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c[j][0]),"+d"(c[j][1]):"d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<ILP;j++) s+=c[j][0]+c[j][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int ILP>
__global__ void dfma_kernel(double* out, const double* in, int iters){
  double a=in[threadIdx.x], b=in[threadIdx.x+32];
  double c[ILP];
  #pragma unroll
  for(int j=0;j<ILP;j++) c[j]=j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<ILP;j++) c[j]=fma(a,c[j],b);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<ILP;j++) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// mixed: per loop ILP DMMAs and NF DFMAs
template<int ILP,int NF>
__global__ void mixed_kernel(double* out, const double* in, int iters){
  double a=in[threadIdx.x], b=in[threadIdx.x+32];
  double c[ILP][2]; double f[NF];
  #pragma unroll
  for(int j=0;j<ILP;j++){c[j][0]=0;c[j][1]=0;}
  #pragma unroll
  for(int j=0;j<NF;j++) f[j]=j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<ILP;j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c[j][0]),"+d"(c[j][1]):"d"(a),"d"(b));
    #pragma unroll
    for(int j=0;j<NF;j++) f[j]=fma(a,f[j],b);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<ILP;j++) s+=c[j][0]+c[j][1];
  #pragma unroll
  for(int j=0;j<NF;j++) s+=f[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); return ms;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int nsm=p.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n",p.name,nsm,p.clockRate);
  double *in,*out; CK(cudaMalloc(&in,4096*8)); CK(cudaMalloc(&out,nsm*32*1024*8));
  CK(cudaMemset(in,0,4096*8));
  const int iters=200000;
  int warps_list[]={1,2,4,8,16,32};
  for(int w: warps_list){
    int threads=32*w; if(threads>1024) continue;
    { float ms=timeit([&]{dmma_kernel<8><<<nsm,threads>>>(out,in,iters);});
      double fl=2.0*256*8*(double)iters*w*nsm; printf("DMMA884 ILP8 warps/SM %2d: %8.3f ms %7.2f TFLOP/s\n",w,ms,fl/ms*1e-9); }
    { float ms=timeit([&]{dmma_kernel<2><<<nsm,threads>>>(out,in,iters);});
      double fl=2.0*256*2*(double)iters*w*nsm; printf("DMMA884 ILP2 warps/SM %2d: %8.3f ms %7.2f TFLOP/s\n",w,ms,fl/ms*1e-9); }
    { float ms=timeit([&]{dfma_kernel<8><<<nsm,threads>>>(out,in,iters);});
      double fl=2.0*32*8*(double)iters*w*nsm; printf("DFMA    ILP8 warps/SM %2d: %8.3f ms %7.2f TFLOP/s\n",w,ms,fl/ms*1e-9); }
    { float ms=timeit([&]{mixed_kernel<4,8><<<nsm,threads>>>(out,in,iters);});
      double fl=(2.0*256*4+2.0*32*8)*(double)iters*w*nsm; printf("MIXED 4dmma+8dfma warps/SM %2d: %8.3f ms %7.2f TFLOP/s\n",w,ms,fl/ms*1e-9); }
  }
  // 2 CTAs per SM variant to test occupancy effects
  { float ms=timeit([&]{dmma_kernel<8><<<nsm*2,256>>>(out,in,iters);});
    double fl=2.0*256*8*(double)iters*8*nsm*2; printf("DMMA884 ILP8 2CTAx8warps: %8.3f ms %7.2f TFLOP/s\n",ms,fl/ms*1e-9); }
  return 0;
}
