// Throw-away microbenchmark: smem/global atomic and FP64 math throughput on B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

constexpr int ROWS=32, STRIDE=257, ITERS=2000;

template<int MODE>
__global__ void __launch_bounds__(256) k_smem(unsigned long long* sink, unsigned long long* gl, int nrowsGlobal){
  extern __shared__ unsigned long long h[];
  unsigned int* h32 = (unsigned int*)h;
  for(int i=threadIdx.x;i<ROWS*STRIDE;i+=blockDim.x) h[i]=0;
  __syncthreads();
  int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  unsigned int s = (blockIdx.x*8+warp)*2654435761u+12345u;
  unsigned long long acc=0;
  for(int it=0;it<ITERS;it++){
    s = s*1664525u+1013904223u;
    int bin = ((s>>8)&127) + ((lane*3)>>3);   // smooth variation across lanes
    #pragma unroll
    for(int e=0;e<4;e++){
      int b = bin + e*13; 
      int idx = lane*STRIDE + b;
      unsigned long long w = (unsigned long long)(s+e) << 13;
      if(MODE==0){ atomicAdd(&h32[2*idx], (unsigned int)w); }
      else if(MODE==1){ unsigned int wl=(unsigned int)w, wh=(unsigned int)(w>>32);
        unsigned int old=atomicAdd(&h32[2*idx], wl); unsigned int c = (old+wl<old);
        atomicAdd(&h32[2*idx+1], wh+c); }
      else if(MODE==2){ atomicAdd(&h[idx], w); }
      else if(MODE==3){ h[idx] += w; }
      else if(MODE==4){ size_t g = ((size_t)((blockIdx.x*8+warp)%nrowsGlobal)*32 + lane)*256 + b; atomicAdd(&gl[g], w); }
      else if(MODE==5){ double* hd=(double*)h; atomicAdd(&hd[idx], (double)w); }
    }
  }
  __syncthreads();
  for(int i=threadIdx.x;i<ROWS*STRIDE;i+=blockDim.x) acc+=h[i];
  if(acc==0x1234567) sink[0]=acc;
}

template<int MODE>
__global__ void __launch_bounds__(256) k_math(double* out, double seed){
  double x0 = seed + threadIdx.x*1e-3 + blockIdx.x*1e-6, x1=x0*1.5, x2=x0*2.5, x3=x0*3.5;
  double a0=0,a1=0,a2=0,a3=0;
  for(int it=0;it<ITERS;it++){
    if(MODE==0){ a0+=sqrt(x0); a1+=sqrt(x1); a2+=sqrt(x2); a3+=sqrt(x3); }
    if(MODE==1){ a0+=log(x0); a1+=log(x1); a2+=log(x2); a3+=log(x3); }
    if(MODE==2){ a0+=1.0/x0; a1+=1.0/x1; a2+=1.0/x2; a3+=1.0/x3; }
    if(MODE==3){ a0=fma(a0,x0,x1); a1=fma(a1,x1,x2); a2=fma(a2,x2,x3); a3=fma(a3,x3,x0); }
    if(MODE==4){ a0+=(double)__logf((float)x0); a1+=(double)__logf((float)x1); a2+=(double)__logf((float)x2); a3+=(double)__logf((float)x3);}
    if(MODE==5){ a0+=asin(x0*0.01); a1+=atan2(x1,x0); a2+=asin(x2*0.01); a3+=atan2(x3,x1); }
    x0+=1e-3; x1+=1e-3; x2+=1e-3; x3+=1e-3;
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3;
}

template<typename F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("dev %s SMs %d clk %d kHz smemPerSM %zu\n", p.name, p.multiProcessorCount, clk, p.sharedMemPerMultiprocessor);
  unsigned long long *sink,*gl; CK(cudaMalloc(&sink,8)); int nrowsG=800; CK(cudaMalloc(&gl,(size_t)nrowsG*32*256*8)); CK(cudaMemset(gl,0,(size_t)nrowsG*32*256*8));
  double* out; CK(cudaMalloc(&out, 148*3*256*8*4));
  size_t smem = ROWS*STRIDE*8;
  int grid = p.multiProcessorCount*3;
  const char* names[]={"ATOMS.ADD.32 (1 per edge)","ATOMS.ADD.32 lo+hi carry","atomicAdd u64 smem (CAS)","plain LDS/STS RMW 64","REDG.ADD.64 global(52MB)","atomicAdd f64 smem (CAS)"};
  #define RUN(M) { cudaFuncSetAttribute(k_smem<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    float ms=timeit([&]{ k_smem<M><<<grid,256,smem>>>(sink,gl,nrowsG); }); CK(cudaGetLastError()); \
    double ops=(double)grid*8*ITERS*4; /* warp-level 64-bit adds */ \
    printf("%-28s %8.3f ms  %7.2f Gwarp-adds/s  => %6.2f ns per warp-add per SM, %6.1f cyc@1.9GHz per warp-add per SM\n", names[M], ms, ops/ms*1e-6, ms*1e6/ops*p.multiProcessorCount, ms*1e6/ops*p.multiProcessorCount*1.9); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
  const char* mn[]={"dsqrt","dlog","drcp","dfma","logf+cvt","asin/atan2"};
  #define RUNM(M) { int g2=p.multiProcessorCount*8; float ms=timeit([&]{ k_math<M><<<g2,256>>>(out,1.0); }); CK(cudaGetLastError()); \
    double ops=(double)g2*256*ITERS*4; printf("%-12s %8.3f ms  %8.2f Gop/s  => %6.2f cyc@1.9GHz per warp-op per SM\n", mn[M], ms, ops/ms*1e-6, ms*1e-3*1.9e9/(ops/32/p.multiProcessorCount)); }
  RUNM(0) RUNM(1) RUNM(2) RUNM(3) RUNM(4) RUNM(5)
  return 0;
}
