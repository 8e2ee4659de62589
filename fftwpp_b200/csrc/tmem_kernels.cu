// tmem_kernels.cu -- fused 1-D convolution rows with Blackwell tensor memory
// (TMEM) as thread-private scratch.
//
// The fused row kernels of fast_kernels.cu keep the transformed second input
// in registers while the first input is transformed, and park the running sum
// over residues in shared memory: 128 registers, 16 warps per SM, and the
// shared-memory pipe -- the busiest unit of the kernel -- also carries the
// parking traffic.  Here both live in TMEM (256 KB per SM, otherwise unused on
// this path: no tensor-core math): each thread stores its 8 complex values to
// its own TMEM lane with tcgen05.st and takes them back with tcgen05.ld when
// the multiplier / the accumulation needs them.  One transform's worth of
// registers is live at a time: 80 registers, 3 CTAs = 24 warps per SM, and no
// parking traffic on the LSU pipe.
//
// Reference loop being replaced: Convolution::convolveRaw residue loop with
// multBinary/multcorrelation (convolve.cc:7513-7575,33-110), fftPad forward1/
// backward1 (convolve.cc:849-958,1482-1546); shape p=1, q=2, L == m == 512.
//
// TMEM rules used: allocation by one warp (power-of-two columns, here 128 per
// CTA), a warp reaches only the lanes of its quadrant 32*(warp%4), shape
// 32x32b = one 32-bit word per thread per column.

#include "regfft.cuh"
#include "warpfft.cuh"

#include <mutex>

#ifndef FFTWPP_TMEM_DEFAULT
#define FFTWPP_TMEM_DEFAULT 0
#endif

namespace fftwpp_gpu {

namespace {

// ---- tensor memory (TMEM) as thread-private scratch ----
__device__ __forceinline__ void tmemAlloc(unsigned *slot, int cols)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"((unsigned) __cvta_generic_to_shared(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmemFree(unsigned taddr, int cols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}
// 4 complex doubles (16 x 32 bit) of this thread's lane at columns [taddr, +16)
__device__ __forceinline__ void tmemSt4(unsigned taddr, const double2 *v)
{
  unsigned r[16];
#pragma unroll
  for(int i=0; i < 4; ++i) {
    r[4*i]=__double2loint(v[i].x); r[4*i+1]=__double2hiint(v[i].x);
    r[4*i+2]=__double2loint(v[i].y); r[4*i+3]=__double2hiint(v[i].y);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),
                  "r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmemLd4(unsigned taddr, double2 *v)
{
  unsigned r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),
                 "=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for(int i=0; i < 4; ++i)
    v[i]=make_double2(__hiloint2double(r[4*i+1],r[4*i]),__hiloint2double(r[4*i+3],r[4*i+2]));
}
__device__ __forceinline__ void tmemWaitSt()
{
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template<int LG, int MULT>
__global__ void __launch_bounds__(256,3)
fast_conv_rows_tm(PlanDev P, const SubBlockDev *__restrict__ sbs,
                  double2 *f0, const double2 *f1, double scale,
                  const double2 zstep, long long nrows, long long rs,
                  int tabid, long long ngroups)
{
  typedef RegFFT<LG> RF;
  const int M=1 << LG;
  const int TPT=M/8;
  const int NT=256;
  const int ROWS=NT/TPT;
  const int BUF=M+M/8;
  const int TWN=RF::twCount();
  extern __shared__ __align__(16) double2 sm[];
  __shared__ unsigned tmemBase;
  double2 *tws=sm;
  double2 *zs=sm+TWN;
  double2 *bufs=zs+TPT;
  const int rowInCta=threadIdx.x/TPT;
  const int tau=threadIdx.x % TPT;
  for(int i=threadIdx.x; i < TWN; i += NT) tws[i]=__ldg(P.tab[tabid].tw8+i);
  for(int i=threadIdx.x; i < TPT; i += NT) zs[i]=zeta(P,modN(P,sbs[1].k0,i));
  const int warp=threadIdx.x >> 5;
  if(warp == 0) tmemAlloc(&tmemBase,128);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // this warp's lane quadrant, 64 columns per warp: [Y 32][acc 32]
  const unsigned tY=tmemBase+(((unsigned) (32*(warp & 3))) << 16)+(unsigned) ((warp >> 2)*64);
  const unsigned tA=tY+32;

  RowLayout lay;
  lay.base=rowInCta*BUF;
  lay.barid=TPT > 32 ? 1+rowInCta : 0;
  lay.nthreads=TPT;

  for(long long grp=blockIdx.x; grp < ngroups; grp += gridDim.x) {
    long long row=grp*ROWS+rowInCta;
    const bool live=row < nrows;
    if(!live) row=nrows-1;
    double2 *g0=f0+row*rs+tau;
    const double2 *g1=f1+row*rs+tau;
    {
      const long long ngrp=grp+gridDim.x;
      if(ngrp < ngroups) {
        long long nrow=ngrp*ROWS+rowInCta;
        if(nrow >= nrows) nrow=nrows-1;
        const char *p0=(const char *) (f0+nrow*rs);
        const char *p1=(const char *) (f1+nrow*rs);
        for(int off=tau*128; off < M*16; off += TPT*128) {
          asm volatile("prefetch.global.L2 [%0];" :: "l"(p0+off));
          asm volatile("prefetch.global.L2 [%0];" :: "l"(p1+off));
        }
      }
    }
#pragma unroll 1
    for(int isb=0; isb < 2; ++isb) {
      const bool hz=isb == 1;
      double2 x[1][8];
      // ---- second input: transform, park the spectrum in tensor memory ----
#pragma unroll
      for(int t=0; t < 8; ++t) x[0][t]=g1[TPT*t];
      if(hz) {
        double2 z=zs[tau];
#pragma unroll
        for(int t=0; t < 8; ++t) {
          x[0][t]=fmul(x[0][t],z);
          if(t < 7) z=fmul(z,zstep);
        }
      }
      RF::template forward<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
      tmemSt4(tY,x[0]);
      tmemSt4(tY+16,x[0]+4);
      // ---- first input ----
#pragma unroll
      for(int t=0; t < 8; ++t) x[0][t]=g0[TPT*t];
      if(hz) {
        double2 z=zs[tau];
#pragma unroll
        for(int t=0; t < 8; ++t) {
          x[0][t]=fmul(x[0][t],z);
          if(t < 7) z=fmul(z,zstep);
        }
      }
      RF::template forward<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
      tmemWaitSt();
#pragma unroll
      for(int h=0; h < 2; ++h) {
        double2 y[4];
        tmemLd4(tY+16*h,y);
#pragma unroll
        for(int t=0; t < 4; ++t)
          x[0][4*h+t]=MULT == FFTWPP_MULT_BINARY ? fmul(x[0][4*h+t],y[t]) : fmulc(x[0][4*h+t],y[t]);
      }
      RF::template adjoint<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
      if(!hz) {
        tmemSt4(tA,x[0]);
        tmemSt4(tA+16,x[0]+4);
      } else {
        tmemWaitSt();
        double2 z=zs[tau];
#pragma unroll
        for(int h=0; h < 2; ++h) {
          double2 a[4];
          tmemLd4(tA+16*h,a);
#pragma unroll
          for(int t=0; t < 4; ++t) {
            double2 v=fmulc(x[0][4*h+t],z)+a[t];
            if(4*h+t < 7) z=fmul(z,zstep);
            if(live) g0[TPT*(4*h+t)]=make_double2(v.x*scale,v.y*scale);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if(warp == 0) tmemFree(tmemBase,128);
}

// One row per warp: 512 = 16 x 16 x 2 transforms (warpfft.cuh), the second
// input's spectrum and the running sum over residues parked in tensor memory.
template<int MULT, int WARPS>
__global__ void __launch_bounds__(32*WARPS,16/WARPS)
fast_conv_rows_wtm(PlanDev P, const SubBlockDev *__restrict__ sbs,
                   double2 *f0, const double2 *f1, double scale,
                   const double2 zstep, long long nrows, long long rs,
                   int tabid)
{
  typedef WarpFFT512 FFT;
  const int NT=32*WARPS;
  const int COLS=WARPS <= 4 ? 128 : (WARPS <= 8 ? 256 : 512);
  extern __shared__ __align__(16) double2 sm[];
  __shared__ unsigned tmemBase;
  double2 *twa=sm;          // omega_512^{k l}, k < 16, l < 32
  double2 *zs=sm+512;       // zeta^{k1 lane}
  double2 *bufs=zs+32;
  const int warp=threadIdx.x >> 5;
  const int lane=threadIdx.x & 31;
  {
    const double2 *om=P.tab[tabid].omega;
    for(int i=threadIdx.x; i < 512; i += NT)
      twa[i]=__ldg(om+(((i >> 5)*(i & 31)) & 511));
    for(int i=threadIdx.x; i < 32; i += NT) zs[i]=zeta(P,modN(P,sbs[1].k0,i));
  }
  if(warp == 0) tmemAlloc(&tmemBase,COLS);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // 128 columns per warp: [spectrum of the second input 64][running sum 64]
  const unsigned tY=tmemBase+(((unsigned) (32*(warp & 3))) << 16)+(unsigned) ((warp >> 2)*128);
  const unsigned tA=tY+64;
  double2 *buf=bufs+warp*FFT::BUF;

  const long long nw=(long long) gridDim.x*WARPS;
  for(long long row=(long long) blockIdx.x*WARPS+warp; row < nrows; row += nw) {
    double2 *g0=f0+row*rs+lane;
    const double2 *g1=f1+row*rs+lane;
    if(row+nw < nrows) {
      const char *p0=(const char *) (f0+(row+nw)*rs);
      const char *p1=(const char *) (f1+(row+nw)*rs);
      for(int off=lane*128; off < 512*16; off += 32*128) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p0+off));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p1+off));
      }
    }
#pragma unroll 1
    for(int isb=0; isb < 2; ++isb) {
      const bool hz=isb == 1;
      double2 x[16];
#pragma unroll
      for(int t=0; t < 16; ++t) x[t]=g1[32*t];
      if(hz) {
        double2 z=zs[lane];
#pragma unroll
        for(int t=0; t < 16; ++t) {
          x[t]=fmul(x[t],z);
          if(t < 15) z=fmul(z,zstep);
        }
      }
      FFT::forward(x,lane,twa,buf);
#pragma unroll
      for(int h=0; h < 4; ++h) tmemSt4(tY+16*h,x+4*h);
#pragma unroll
      for(int t=0; t < 16; ++t) x[t]=g0[32*t];
      if(hz) {
        double2 z=zs[lane];
#pragma unroll
        for(int t=0; t < 16; ++t) {
          x[t]=fmul(x[t],z);
          if(t < 15) z=fmul(z,zstep);
        }
      }
      FFT::forward(x,lane,twa,buf);
      tmemWaitSt();
#pragma unroll
      for(int h=0; h < 4; ++h) {
        double2 y[4];
        tmemLd4(tY+16*h,y);
#pragma unroll
        for(int t=0; t < 4; ++t)
          x[4*h+t]=MULT == FFTWPP_MULT_BINARY ? fmul(x[4*h+t],y[t]) : fmulc(x[4*h+t],y[t]);
      }
      FFT::adjoint(x,lane,twa,buf);
      if(!hz) {
#pragma unroll
        for(int h=0; h < 4; ++h) tmemSt4(tA+16*h,x+4*h);
      } else {
        tmemWaitSt();
        double2 z=zs[lane];
#pragma unroll
        for(int h=0; h < 4; ++h) {
          double2 a[4];
          tmemLd4(tA+16*h,a);
#pragma unroll
          for(int t=0; t < 4; ++t) {
            double2 v=fmulc(x[4*h+t],z)+a[t];
            if(4*h+t < 15) z=fmul(z,zstep);
            g0[32*(4*h+t)]=make_double2(v.x*scale,v.y*scale);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if(warp == 0) tmemFree(tmemBase,COLS);
}

// FFTWPP_CONV_TMEM: 0 = off (register/shared-memory kernels of
// fast_kernels.cu), 1 = radix-8 rows with TMEM parking (24 warps/SM),
// 2 = one-warp rows with TMEM parking (16 rows in flight per SM)
int tmemMode()
{
  static int mode=-1;
  if(mode < 0) {
    const char *s=getenv("FFTWPP_CONV_TMEM");
    mode=s ? atoi(s) : FFTWPP_TMEM_DEFAULT;
    if(mode < 0 || mode > 2) mode=0;
  }
  return mode;
}

// zeta_N^{32 k0}: the step between a lane's successive points (one-warp rows)
double2 zstepw(Plan *pl)
{
  const long double ang=2.0L*3.141592653589793238462643383279502884L*
    (long double) ((pl->hsub[1].k0*32ull) % (unsigned long long) pl->dev.N)/
    (long double) pl->dev.N;
  return make_double2((double) cosl(ang),(double) sinl(ang));
}

template<class K>
int allowSmemT(K kernel, size_t bytes)
{
  static std::mutex mu;
  static std::vector<std::pair<const void *,int> > done;
  int dev=0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for(size_t i=0; i < done.size(); ++i)
    if(done[i].first == (const void *) kernel && done[i].second == dev)
      return 0;
  cudaError_t e=cudaFuncSetAttribute(kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int) bytes);
  if(e != cudaSuccess) return cuda_fail(e,"cudaFuncSetAttribute");
  done.push_back(std::make_pair((const void *) kernel,dev));
  return 0;
}

} // namespace

// Same contract as fast_try_convolve: 1 handled, 0 not applicable, <0 error.
int tmem_try_convolve(Plan *pl, void *const *f, uint32_t A, uint32_t B,
                      int mult, double scale, uint64_t nrows, uint64_t rs,
                      cudaStream_t st)
{
  FastInfo *fi=pl->fast;
  const int mode=tmemMode();
  if(mode == 0 || !fi || !fi->uniform || fi->nterm != 1) return 0;
  const PlanDev& d=pl->dev;
  if(fi->log2m != 9 || d.kind != FFTWPP_KIND_COMPLEX || d.C != 1 || d.S != 1)
    return 0;
  if(A != 2 || B != 1) return 0;
  if(mult != FFTWPP_MULT_BINARY && mult != FFTWPP_MULT_CORRELATION) return 0;
  const int M=512, TPT=64, ROWS=4, BUF=M+M/8;
  if(pl->hsub.size() != 2 || pl->hsub[0].k0 != 0 || pl->hsub[1].k0 == 0 ||
     d.jmax != M || d.jmin != 0)
    return 0;
  int tabid=-1;
  for(int k=0; k < 2; ++k)
    if(d.tab[k].n == M) tabid=k;
  if(tabid < 0) return 0;
  const uint64_t ngroups=(nrows+ROWS-1)/ROWS;
  if(ngroups == 0) return 1;
  int sms=148, dev=0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms,cudaDevAttrMultiProcessorCount,dev);
  const uint64_t grid=std::min<uint64_t>(ngroups,(uint64_t) sms*3);
  const size_t smem=((size_t) RegFFT<9>::twCount()+TPT+(size_t) ROWS*BUF)*
    sizeof(double2);
  const long double ang=2.0L*3.141592653589793238462643383279502884L*
    (long double) ((pl->hsub[1].k0*(unsigned long long) TPT) %
                   (unsigned long long) d.N)/(long double) d.N;
  const double2 zstep=make_double2((double) cosl(ang),(double) sinl(ang));
  int rc=0;
  if(mode == 2) {
    const int W=8;
    const uint64_t gridw=std::min<uint64_t>((nrows+W-1)/W,(uint64_t) sms*2);
    const size_t smw=(512+32+(size_t) W*WarpFFT512::BUF)*sizeof(double2);
    if(mult == FFTWPP_MULT_BINARY) {
      rc=allowSmemT(fast_conv_rows_wtm<FFTWPP_MULT_BINARY,W>,smw);
      if(rc) return rc;
      prof_begin(4*pl->tag+2,st);
      fast_conv_rows_wtm<FFTWPP_MULT_BINARY,W><<<(unsigned) gridw,32*W,smw,st>>>
        (pl->dev,pl->dsub,(double2 *) f[0],(const double2 *) f[1],scale,
         zstepw(pl),(long long) nrows,(long long) rs,tabid);
    } else {
      rc=allowSmemT(fast_conv_rows_wtm<FFTWPP_MULT_CORRELATION,W>,smw);
      if(rc) return rc;
      prof_begin(4*pl->tag+2,st);
      fast_conv_rows_wtm<FFTWPP_MULT_CORRELATION,W>
        <<<(unsigned) gridw,32*W,smw,st>>>
        (pl->dev,pl->dsub,(double2 *) f[0],(const double2 *) f[1],scale,
         zstepw(pl),(long long) nrows,(long long) rs,tabid);
    }
    rc=check_launch("fast_conv_rows_wtm",st);
    return rc ? rc : 1;
  }
  if(mult == FFTWPP_MULT_BINARY) {
    rc=allowSmemT(fast_conv_rows_tm<9,FFTWPP_MULT_BINARY>,smem);
    if(rc) return rc;
    prof_begin(4*pl->tag+2,st);
    fast_conv_rows_tm<9,FFTWPP_MULT_BINARY><<<(unsigned) grid,256,smem,st>>>
      (pl->dev,pl->dsub,(double2 *) f[0],(const double2 *) f[1],scale,zstep,
       (long long) nrows,(long long) rs,tabid,(long long) ngroups);
  } else {
    rc=allowSmemT(fast_conv_rows_tm<9,FFTWPP_MULT_CORRELATION>,smem);
    if(rc) return rc;
    prof_begin(4*pl->tag+2,st);
    fast_conv_rows_tm<9,FFTWPP_MULT_CORRELATION><<<(unsigned) grid,256,smem,st>>>
      (pl->dev,pl->dsub,(double2 *) f[0],(const double2 *) f[1],scale,zstep,
       (long long) nrows,(long long) rs,tabid,(long long) ngroups);
  }
  rc=check_launch("fast_conv_rows_tm",st);
  return rc ? rc : 1;
}

} // namespace fftwpp_gpu
