// Calibration micro-benchmarks for the B200 FFT kernels: FP64 issue rate and
// latency, shared-memory LDS/STS.128 and SHFL throughput, per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

template<int ILP, bool FMA>
__global__ void fp64k(double *out, int iters, double a, double b)
{
  double v[ILP];
#pragma unroll
  for(int i=0; i < ILP; ++i) v[i]=threadIdx.x+i;
  for(int it=0; it < iters; ++it) {
#pragma unroll
    for(int i=0; i < ILP; ++i) v[i]=FMA ? fma(v[i],a,b) : v[i]+a;
  }
  double s=0;
#pragma unroll
  for(int i=0; i < ILP; ++i) s += v[i];
  if(s == 12345.678) out[0]=s;
}

template<int MODE> // 0: LDS.128, 1: STS.128, 2: SHFL.32 x4, 3: LDS.64, 4: LDS.128 stride-9 (padded FFT pattern)
__global__ void smemk(double *out, int iters)
{
  extern __shared__ double2 sm[];
  const int n=blockDim.x;
  for(int i=threadIdx.x; i < 2*n; i += n) sm[i]=make_double2(i,1.0);
  __syncthreads();
  double ax=0.0, ay=1.0;
  unsigned base=(unsigned) __cvta_generic_to_shared(sm);
  const int idx=threadIdx.x;
  for(int it=0; it < iters; ++it) {
#pragma unroll
    for(int u=0; u < 8; ++u) {
      int j=(idx+u*32) & (n-1);
      if(MODE == 4) j=((idx*9+u) & (n-1));
      unsigned a=base+16u*j;
      if(MODE == 0 || MODE == 4) {
        double vx,vy;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(a));
        ax += vx; ay += vy;
      } else if(MODE == 1) {
        asm volatile("st.shared.v2.f64 [%0], {%1,%2};" :: "r"(a), "d"(ax), "d"(ay) : "memory");
      } else if(MODE == 2) {
        unsigned lo=__double2loint(ax), hi=__double2hiint(ax), lo2=__double2loint(ay), hi2=__double2hiint(ay);
        asm volatile("shfl.sync.bfly.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(lo) : "r"(1+u));
        asm volatile("shfl.sync.bfly.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(hi) : "r"(1+u));
        asm volatile("shfl.sync.bfly.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(lo2) : "r"(1+u));
        asm volatile("shfl.sync.bfly.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(hi2) : "r"(1+u));
        ax=__hiloint2double(hi,lo); ay=__hiloint2double(hi2,lo2);
      } else {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base+8u*j));
        ax += v;
      }
    }
  }
  if(ax == 12345.678) out[0]=ax+ay;
}

int main()
{
  double *d;
  CK(cudaMalloc(&d,1024));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p,0));
  int clk=0;
  cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,0);
  printf("%s SMs %d clock %d kHz\n",p.name,p.multiProcessorCount,clk);
  cudaEvent_t e0,e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int SM=p.multiProcessorCount;
  const int iters=20000;
#define RUNF(ILP,FMA,WARPS) { \
    fp64k<ILP,FMA><<<SM,32*WARPS>>>(d,100,1.0000001,1e-9); \
    cudaEventRecord(e0); fp64k<ILP,FMA><<<SM,32*WARPS>>>(d,iters,1.0000001,1e-9); cudaEventRecord(e1); \
    CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); \
    double ops=(double) iters*ILP*32*WARPS; \
    printf("fp64 %s ilp %d warps/SM %2d: %.1f lane-ops/clk/SM (at %d kHz), %.2f ms\n",FMA?"DFMA":"DADD",ILP,WARPS, \
           ops/(ms*1e-3)/(clk*1e3),clk,ms); }
  RUNF(1,true,1) RUNF(2,true,1) RUNF(4,true,1) RUNF(8,true,1)
  RUNF(1,true,4) RUNF(4,true,4) RUNF(8,true,4) RUNF(8,true,8) RUNF(8,true,16) RUNF(4,true,32)
  RUNF(8,false,4) RUNF(8,false,16)
#define RUNS(MODE,NAME,WARPS,BYTES) { \
    smemk<MODE><<<SM,32*WARPS,2*32*WARPS*16>>>(d,10); \
    cudaEventRecord(e0); smemk<MODE><<<SM,32*WARPS,2*32*WARPS*16>>>(d,iters); cudaEventRecord(e1); \
    CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); \
    double ins=(double) iters*8*WARPS; \
    printf("%s warps/SM %2d: %.3f warp-instr/clk/SM = %.1f B/clk/SM, %.2f ms\n",NAME,WARPS, \
           ins/(ms*1e-3)/(clk*1e3),ins*BYTES/(ms*1e-3)/(clk*1e3),ms); }
  RUNS(0,"LDS.128",4,512) RUNS(0,"LDS.128",16,512) RUNS(0,"LDS.128",32,512)
  RUNS(1,"STS.128",4,512) RUNS(1,"STS.128",16,512)
  RUNS(3,"LDS.64 ",16,256)
  RUNS(4,"LDS.128 stride9",16,512)
  RUNS(2,"SHFLx4 ",4,512) RUNS(2,"SHFLx4 ",16,512) RUNS(2,"SHFLx4 ",32,512)
  return 0;
}
