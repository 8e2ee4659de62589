// tmem_kernels.cu -- fused 1-D convolution rows with Blackwell tensor memory
// (TMEM) as thread-private scratch.
//
// The fused row kernels of fast_kernels.cu keep the transformed second input
// in registers while the first input is transformed, and park the running sum
// over residues in shared memory.  Here both live in TMEM (256 KB per SM,
// otherwise unused on this path: no tensor-core math): each thread stores its
// complex values to its own TMEM lane with tcgen05.st and takes them back with
// tcgen05.ld when the multiplier / the accumulation needs them.
//
//   fast_conv_rows_long  rows of m = 8192 and m = 4096 points (default for
//                        those shapes): 16 points per thread; without TMEM
//                        the spectrum and the sum would not fit on chip
//   fast_conv_rows_tm    m = 512, radix-8 rows, 80 registers, 24 warps/SM
//   fast_conv_rows_wtm   m = 512, one-warp rows (warpfft.cuh)
//                        (both opt-in, FFTWPP_CONV_TMEM=1/2: measured slower
//                        than fast_conv_rows_q2, kept as documented experiments)
//
// Reference loop being replaced: Convolution::convolveRaw residue loop with
// multBinary/multcorrelation (convolve.cc:7513-7575,33-110), fftPad forward1/
// backward1 (convolve.cc:849-958,1482-1546); shape p=1, L <= m.
//
// TMEM rules used: allocation by one warp (power-of-two columns), a warp
// reaches only the lanes of its quadrant 32*(warp%4), shape 32x32b = one
// 32-bit word per thread per column (tmem.cuh).

#include "regfft.cuh"
#include "warpfft.cuh"
#include "tmem.cuh"

#include <mutex>

// smallest log2(m) served by fast_conv_rows_long.  11 also routes 2048-point
// rows to it: measured 0.920 vs 0.992 ms for 16384 rows against
// fast_conv_rows_q2 (profiles/README.md); not the default, because no
// BASELINE shape uses such rows and only part of the suite ran on that build.
#ifndef FFTWPP_LONG_MIN_LG
#define FFTWPP_LONG_MIN_LG 12
#endif

#ifndef FFTWPP_TMEM_DEFAULT
#define FFTWPP_TMEM_DEFAULT 0
#endif

namespace fftwpp_gpu {

namespace {

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

// ---------------------------------------------------------------------------
// Long rows: m = 8192 and m = 4096, one CTA per row, fully fused
// ---------------------------------------------------------------------------
// A row of 8192 complex words is 128 KB: the padded exchange buffer of ONE
// transform fills most of an SM's shared memory, so the register/shared-memory
// kernels stop at m = 4096 and longer rows used to take the two-stage path
// (three HBM round trips).  Here every thread owns 16 complex points (64
// registers), and everything that must survive a transform lives in tensor
// memory: the transformed second input (64 columns per warp) and the running
// sum over residues (64 more) -- for m = 8192, 4 warps per lane quadrant x 128
// columns = all 512 columns, 256 KB, of the SM's TMEM; for m = 4096 two CTAs
// of 256 columns each.  HBM traffic is the algorithmic minimum (read both
// inputs, write one output; the second residue re-reads its inputs from L2 --
// an L2 prefetch of the next row evicts them and measured 7 % slower).
//
// Reference loop being replaced: Convolution::convolveRaw over the residues
// with multBinary/multcorrelation (convolve.cc:7513-7575,33-110) on
// fftPad::forward1/backward1 (convolve.cc:849-958,1482-1546), p=1, L <= m.

// m = 16 x SUB (8192 = 16 x 512, 4096 = 16 x 256, 2048 = 16 x 128), SUB threads.  A thread's 16
// points are SUB apart, so the first pass is a radix-16 butterfly in
// registers; ONE CTA-wide exchange then hands every warp whole sub-transforms
// (8192: one of 512 points, two virtual threads per lane; 4096: two of 256
// points, one virtual thread per lane each), whose remaining exchanges need
// only __syncwarp().  Compared with RegFFT<LG> over the whole row: one
// exchange less for 8192, and two CTA barriers per transform instead of six or
// eight, so the warps drift apart and overlap butterflies with shared-memory
// traffic.  The order of the transformed points differs from RegFFT<LG>'s; the
// fused convolution only multiplies pointwise, so any order shared by
// forward() and adjoint() serves.
template<int LGV>
struct LongRow16 {
  static const int LG=LGV;
  static const int SUBLG=LG-4;
  typedef RegFFT<SUBLG> RS;
  static const int M=1 << LG;
  static const int SUB=1 << SUBLG;       // points of a sub-transform
  static const int NT=SUB;               // threads
  static const int VT=2;
  static const int TPT=NT*VT;
  static const int NW=NT/32;             // warps
  static const int TS=SUB/8;             // virtual threads of a sub-transform
  static const int SUBBUF=SUB+SUB/8;
  static const int BUF=16*SUBBUF;
  static const int NR8=RS::NR8;
  static_assert(TS == 64 || TS == 32 || TS == 16,
                "a sub-transform must live inside one warp");

  // tables: w_M^j, j < SUB | per radix-8 pass i of the sub-transform (legs
  // 2^ls_i apart, ls_i > 0): w_SUB^{j 8^i}, j < 2^ls_i
  static __host__ __device__ constexpr int wOff(int i) {
    int off=SUB;
    for(int k=0; k < i; ++k) off += 1 << (SUBLG-3*(k+1));
    return off;
  }
  static constexpr int W1N=(wOff(NR8)+15) & ~15;

  static __device__ __forceinline__ int pad(int p) {return p+(p >> 3);}
  // sub-transform and virtual thread served by array a of this thread
  // (TS < 32: 32/TS sub-transforms per warp and array, side by side)
  static __device__ __forceinline__ int ksub(int warp, int lane, int a) {
    return TS == 64 ? warp : TS == 32 ? warp+NW*a :
      (2*warp+a)*(32/TS)+lane/TS;
  }
  static __device__ __forceinline__ int tau(int lane, int a) {
    return TS == 64 ? lane+32*a : TS == 32 ? lane : lane % TS;
  }

  static __device__ __forceinline__ void init(const double2 *tw, double2 *w1s,
                                              int tid)
  {
    // the u=1 block of pass 0 of the tw8 table holds w_M^j, j < M/8
    for(int j=tid; j < SUB; j += NT) w1s[j]=__ldg(tw+j);
#pragma unroll
    for(int i=0; i < NR8; ++i) {
      const int ls=SUBLG-3*(i+1);
      if(ls > 0)
        for(int j=tid; j < (1 << ls); j += NT)
          w1s[wOff(i)+j]=__ldg(tw+((16*j) << (3*i)));
    }
  }

  template<bool CONJ>
  static __device__ __forceinline__ double2 cm(double2 a, double2 w)
  {
    return CONJ ? fmulc(a,w) : fmul(a,w);
  }

  // x[u] *= w^u (CONJ: conj(w)^u)
  template<bool CONJ>
  static __device__ __forceinline__ void twiddle(double2 (&x)[8], double2 w1)
  {
    x[1]=cm<CONJ>(x[1],w1);
    const double2 w2=fmul(w1,w1);
    x[2]=cm<CONJ>(x[2],w2);
    const double2 w3=fmul(w1,w2);
    x[3]=cm<CONJ>(x[3],w3);
    const double2 w4=fmul(w2,w2);
    x[4]=cm<CONJ>(x[4],w4);
    x[5]=cm<CONJ>(x[5],fmul(w1,w4));
    x[6]=cm<CONJ>(x[6],fmul(w3,w3));
    x[7]=cm<CONJ>(x[7],fmul(w3,w4));
  }

  // y_k *= w^k (CONJ: conj), k = k8+8a over x[a][k8]
  template<bool CONJ>
  static __device__ __forceinline__ void twiddle16(double2 (&x)[VT][8],
                                                   double2 w1)
  {
    const double2 w2=fmul(w1,w1);
    const double2 w4=fmul(w2,w2);
    const double2 w8=fmul(w4,w4);
    x[1][0]=cm<CONJ>(x[1][0],w8);
    x[0][1]=cm<CONJ>(x[0][1],w1);
    x[1][1]=cm<CONJ>(x[1][1],fmul(w1,w8));
    x[0][2]=cm<CONJ>(x[0][2],w2);
    x[1][2]=cm<CONJ>(x[1][2],fmul(w2,w8));
    const double2 w3=fmul(w1,w2);
    x[0][3]=cm<CONJ>(x[0][3],w3);
    x[1][3]=cm<CONJ>(x[1][3],fmul(w3,w8));
    x[0][4]=cm<CONJ>(x[0][4],w4);
    x[1][4]=cm<CONJ>(x[1][4],fmul(w4,w8));
    const double2 w5=fmul(w1,w4);
    x[0][5]=cm<CONJ>(x[0][5],w5);
    x[1][5]=cm<CONJ>(x[1][5],fmul(w5,w8));
    const double2 w6=fmul(w3,w3);
    x[0][6]=cm<CONJ>(x[0][6],w6);
    x[1][6]=cm<CONJ>(x[1][6],fmul(w6,w8));
    const double2 w7=fmul(w3,w4);
    x[0][7]=cm<CONJ>(x[0][7],w7);
    x[1][7]=cm<CONJ>(x[1][7],fmul(w7,w8));
  }

  // z[k] *= exp(2 pi i k/16) (CONJ: its conjugate)
  template<bool CONJ>
  static __device__ __forceinline__ void rot16(double2 (&z)[8])
  {
    const double c=0.92387953251128675613; // cos(pi/8)
    const double s=0.38268343236508977173; // sin(pi/8)
    const double h=0.70710678118654752440;
    z[1]=cm<CONJ>(z[1],make_double2(c,s));
    z[2]=cm<CONJ>(z[2],make_double2(h,h));
    z[3]=cm<CONJ>(z[3],make_double2(s,c));
    z[4]=CONJ ? rot<-1>(z[4]) : rot<1>(z[4]);
    z[5]=cm<CONJ>(z[5],make_double2(-s,c));
    z[6]=cm<CONJ>(z[6],make_double2(-h,h));
    z[7]=cm<CONJ>(z[7],make_double2(-c,s));
  }

  // exchange inside the warp's sub-transform regions
  static __device__ __forceinline__ void warpExchange(double2 (&x)[VT][8],
                                                      int warp, int lane,
                                                      int lsFrom, int lsTo,
                                                      double2 *buf)
  {
    __syncwarp();
#pragma unroll
    for(int a=0; a < VT; ++a)
#pragma unroll
      for(int t=0; t < 8; ++t)
        buf[ksub(warp,lane,a)*SUBBUF+pad(RS::pos(tau(lane,a),t,lsFrom))]=x[a][t];
    __syncwarp();
#pragma unroll
    for(int a=0; a < VT; ++a)
#pragma unroll
      for(int t=0; t < 8; ++t)
        x[a][t]=buf[ksub(warp,lane,a)*SUBBUF+pad(RS::pos(tau(lane,a),t,lsTo))];
  }

  template<int SIGN>
  static __device__ __forceinline__ void remainder(double2 (&x)[VT][8])
  {
#pragma unroll
    for(int a=0; a < VT; ++a) {
      if(RS::REM == 2) {
        bfly4<SIGN>(x[a][0],x[a][1],x[a][2],x[a][3]);
        bfly4<SIGN>(x[a][4],x[a][5],x[a][6],x[a][7]);
      } else if(RS::REM == 1) {
        bfly2(x[a][0],x[a][1]);
        bfly2(x[a][2],x[a][3]);
        bfly2(x[a][4],x[a][5]);
        bfly2(x[a][6],x[a][7]);
      }
    }
  }

  // in: x[a][t]=W[tid+NT*(a+2t)]; out: some fixed order of the transform
  static __device__ __forceinline__ void forward(double2 (&x)[VT][8], int tid,
                                                 const double2 *w1s,
                                                 double2 *buf)
  {
    const int warp=tid >> 5, lane=tid & 31;
    // radix 16 over s=a+2t: y_k = sum_s x_s w_16^{sk}, k=k8+8a
    bfly8<1>(x[0]);
    bfly8<1>(x[1]);
    rot16<false>(x[1]);
#pragma unroll
    for(int k=0; k < 8; ++k) bfly2(x[0][k],x[1][k]);
    twiddle16<false>(x,w1s[tid]);
    // y_k[tid] -> sub-transform k, column tid
    __syncthreads();
#pragma unroll
    for(int a=0; a < VT; ++a)
#pragma unroll
      for(int k=0; k < 8; ++k)
        buf[(k+8*a)*SUBBUF+pad(tid)]=x[a][k];
    __syncthreads();
#pragma unroll
    for(int a=0; a < VT; ++a)
#pragma unroll
      for(int t=0; t < 8; ++t)
        x[a][t]=buf[ksub(warp,lane,a)*SUBBUF+pad(tau(lane,a)+TS*t)];
    // the sub-transforms: RegFFT<SUBLG> inside the warp
#pragma unroll
    for(int i=0; i < NR8; ++i) {
      const int ls=SUBLG-3*(i+1);
#pragma unroll
      for(int a=0; a < VT; ++a) {
        bfly8<1>(x[a]);
        if(ls > 0)
          twiddle<false>(x[a],w1s[wOff(i)+(tau(lane,a) & ((1 << ls)-1))]);
      }
      const int lsNext=(i+1 < NR8) ? SUBLG-3*(i+2) : 0;
      if(i+1 < NR8 || RS::REM > 0) warpExchange(x,warp,lane,ls,lsNext,buf);
    }
    remainder<1>(x);
  }

  // exact adjoint of forward()
  static __device__ __forceinline__ void adjoint(double2 (&x)[VT][8], int tid,
                                                 const double2 *w1s,
                                                 double2 *buf)
  {
    const int warp=tid >> 5, lane=tid & 31;
    remainder<-1>(x);
#pragma unroll
    for(int i=NR8-1; i >= 0; --i) {
      const int ls=SUBLG-3*(i+1);
      const int lsPrev=(i+1 < NR8) ? SUBLG-3*(i+2) : 0;
      if(i+1 < NR8 || RS::REM > 0) warpExchange(x,warp,lane,lsPrev,ls,buf);
#pragma unroll
      for(int a=0; a < VT; ++a) {
        if(ls > 0)
          twiddle<true>(x[a],w1s[wOff(i)+(tau(lane,a) & ((1 << ls)-1))]);
        bfly8<-1>(x[a]);
      }
    }
    __syncwarp();
#pragma unroll
    for(int a=0; a < VT; ++a)
#pragma unroll
      for(int t=0; t < 8; ++t)
        buf[ksub(warp,lane,a)*SUBBUF+pad(tau(lane,a)+TS*t)]=x[a][t];
    __syncthreads();
#pragma unroll
    for(int a=0; a < VT; ++a)
#pragma unroll
      for(int k=0; k < 8; ++k)
        x[a][k]=buf[(k+8*a)*SUBBUF+pad(tid)];
    twiddle16<true>(x,w1s[tid]);
#pragma unroll
    for(int k=0; k < 8; ++k) bfly2(x[0][k],x[1][k]);
    rot16<true>(x[1]);
    bfly8<-1>(x[0]);
    bfly8<-1>(x[1]);
  }

  // x[a][t] *= zeta_N^{k0 j} (CONJ: its conjugate), j=tid+NT*a+TPT*t
  template<bool CONJ>
  static __device__ __forceinline__ void residue(double2 (&x)[VT][8],
                                                 const PlanDev& P,
                                                 long long k0, int tid)
  {
    const double2 zst=zeta(P,modN(P,k0,TPT));
#pragma unroll
    for(int a=0; a < VT; ++a) {
      double2 z=zeta(P,modN(P,k0,tid+NT*a));
#pragma unroll
      for(int t=0; t < 8; ++t) {
        x[a][t]=cm<CONJ>(x[a][t],z);
        if(t < 7) z=fmul(z,zst);
      }
    }
  }
};

template<class LR, int MULT>
__global__ void __launch_bounds__(LR::NT,512/LR::NT)
fast_conv_rows_long(PlanDev P, const SubBlockDev *__restrict__ sbs, int nsb,
                    double2 *f0, const double2 *f1, double scale,
                    long long nrows, long long rs, int tabid)
{
  const int TPT=LR::TPT;
  const int NT=LR::NT;
  const int VT=LR::VT;
  const int COLS=32*VT;   // TMEM columns of one parked set per warp
  // of the CTA: NT/128 (at least one) warps per lane quadrant
  const int TCOLS=(NT >= 128 ? NT/128 : 1)*2*COLS;
  extern __shared__ __align__(16) double2 sm[];
  __shared__ unsigned tmemBase;
  double2 *w1s=sm;
  double2 *buf=sm+LR::W1N;
  const int tid=threadIdx.x;
  const int L=P.jmax;

  LR::init(P.tab[tabid].tw8,w1s,tid);
  const int warp=tid >> 5;
  if(warp == 0) tmemAlloc(&tmemBase,TCOLS);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // this warp's lane quadrant; 2*COLS columns per warp: [spectrum | sum]
  const unsigned tY=tmemBase+(((unsigned) (32*(warp & 3))) << 16)+
    (unsigned) ((warp >> 2)*2*COLS);
  const unsigned tA=tY+COLS;

  for(long long row=blockIdx.x; row < nrows; row += gridDim.x) {
    double2 *g0=f0+row*rs;
    const double2 *g1=f1+row*rs;
#pragma unroll 1
    for(int isb=0; isb < nsb; ++isb) {
      const long long k0=sbs[isb].k0;
      // residue twiddles zeta_N^{k0 j}, j=tau_a+TPT*t: base times powers of
      // the step zeta_N^{k0 TPT}
      // (looked up at each use: registers are the scarce resource here)
      double2 x[VT][8];
      // ---- second input: transform, park the spectrum in tensor memory ----
#pragma unroll
      for(int a=0; a < VT; ++a)
#pragma unroll
        for(int t=0; t < 8; ++t) {
          const int j=tid+NT*a+TPT*t;
          x[a][t]=j < L ? g1[j] : make_double2(0.0,0.0);
        }
      if(k0 != 0) LR::template residue<false>(x,P,k0,tid);
      LR::forward(x,tid,w1s,buf);
#pragma unroll
      for(int a=0; a < VT; ++a) {
        tmemSt4(tY+32*a,&x[a][0]);
        tmemSt4(tY+32*a+16,&x[a][4]);
      }
      // ---- first input ----
#pragma unroll
      for(int a=0; a < VT; ++a)
#pragma unroll
        for(int t=0; t < 8; ++t) {
          const int j=tid+NT*a+TPT*t;
          x[a][t]=j < L ? g0[j] : make_double2(0.0,0.0);
        }
      if(k0 != 0) LR::template residue<false>(x,P,k0,tid);
      LR::forward(x,tid,w1s,buf);
      tmemWaitSt();
      // ---- multiplier, inverse transform ----
#pragma unroll
      for(int a=0; a < VT; ++a)
#pragma unroll
        for(int h=0; h < 2; ++h) {
          double2 y[4];
          tmemLd4(tY+32*a+16*h,y);
#pragma unroll
          for(int t=0; t < 4; ++t)
            x[a][4*h+t]=MULT == FFTWPP_MULT_BINARY ? fmul(x[a][4*h+t],y[t]) :
              fmulc(x[a][4*h+t],y[t]);
        }
      LR::adjoint(x,tid,w1s,buf);
      if(k0 != 0) LR::template residue<true>(x,P,k0,tid);
      // ---- running sum over the residues (tensor memory) ----
      if(isb > 0) {
#pragma unroll
        for(int a=0; a < VT; ++a)
#pragma unroll
          for(int h=0; h < 2; ++h) {
            double2 y[4];
            tmemLd4(tA+32*a+16*h,y);
#pragma unroll
            for(int t=0; t < 4; ++t) x[a][4*h+t]=x[a][4*h+t]+y[t];
          }
      }
      if(isb+1 < nsb) {
#pragma unroll
        for(int a=0; a < VT; ++a) {
          tmemSt4(tA+32*a,&x[a][0]);
          tmemSt4(tA+32*a+16,&x[a][4]);
        }
        tmemWaitSt();
      } else {
#pragma unroll
        for(int a=0; a < VT; ++a)
#pragma unroll
          for(int t=0; t < 8; ++t) {
            const int j=tid+NT*a+TPT*t;
            if(j < L)
              g0[j]=make_double2(x[a][t].x*scale,x[a][t].y*scale);
          }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if(warp == 0) tmemFree(tmemBase,TCOLS);
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

// Rows of m = 8192 or 4096 points, p = 1 (L <= m), uniform sub-blocks.
template<class LR>
int launchLongRows(Plan *pl, void *const *f, int mult, double scale,
                   uint64_t nrows, uint64_t rs, cudaStream_t st)
{
  const PlanDev& d=pl->dev;
  int tabid=-1;
  for(int k=0; k < 2; ++k)
    if(d.tab[k].n == LR::M && d.tab[k].tw8) tabid=k;
  if(tabid < 0) return 0;
  if(nrows == 0) return 1;
  const int sms=sm_count();
  const uint64_t grid=std::min<uint64_t>(nrows,(uint64_t) sms*(512/LR::NT));
  const size_t smem=((size_t) LR::W1N+LR::BUF)*sizeof(double2);
  int rc=0;
  if(mult == FFTWPP_MULT_BINARY) {
    rc=allowSmemT(fast_conv_rows_long<LR,FFTWPP_MULT_BINARY>,smem);
    if(rc) return rc;
    prof_begin(4*pl->tag+2,st);
    fast_conv_rows_long<LR,FFTWPP_MULT_BINARY><<<(unsigned) grid,LR::NT,smem,st>>>
      (pl->dev,pl->dsub,(int) pl->hsub.size(),(double2 *) f[0],
       (const double2 *) f[1],scale,(long long) nrows,(long long) rs,tabid);
  } else {
    rc=allowSmemT(fast_conv_rows_long<LR,FFTWPP_MULT_CORRELATION>,smem);
    if(rc) return rc;
    prof_begin(4*pl->tag+2,st);
    fast_conv_rows_long<LR,FFTWPP_MULT_CORRELATION>
      <<<(unsigned) grid,LR::NT,smem,st>>>
      (pl->dev,pl->dsub,(int) pl->hsub.size(),(double2 *) f[0],
       (const double2 *) f[1],scale,(long long) nrows,(long long) rs,tabid);
  }
  rc=check_launch("fast_conv_rows_long",st);
  return rc ? rc : 1;
}

int tryLongRows(Plan *pl, void *const *f, uint32_t A, uint32_t B, int mult,
                double scale, uint64_t nrows, uint64_t rs, cudaStream_t st)
{
  const PlanDev& d=pl->dev;
  const unsigned M=pl->mmax;
  if((M != 8192 && M != 4096 && M != 2048) || d.kind != FFTWPP_KIND_COMPLEX ||
     d.C != 1 ||
     d.S != 1 || A != 2 || B != 1)
    return 0;
  if(mult != FFTWPP_MULT_BINARY && mult != FFTWPP_MULT_CORRELATION) return 0;
  // one or two sub-blocks (explicit rows, p=1 q=2): the shapes the GPU suite
  // covers; the kernel's residue loop is general, but longer lists (q > 2)
  // stay on the kernels of fast_kernels.cu until they are tested here
  if(d.jmin != 0 || d.jmax > (int) M || pl->hsub.empty() ||
     pl->hsub.size() > 2)
    return 0;
  for(size_t i=0; i < pl->hsub.size(); ++i)
    if(pl->hsub[i].mlen != M || pl->hsub[i].nout != M ||
       pl->hsub[i].flags != 0)
      return 0;
  if(M == 8192)
    return launchLongRows<LongRow16<13> >(pl,f,mult,scale,nrows,rs,st);
  if(M == 4096)
    return launchLongRows<LongRow16<12> >(pl,f,mult,scale,nrows,rs,st);
#if FFTWPP_LONG_MIN_LG <= 11
  return launchLongRows<LongRow16<11> >(pl,f,mult,scale,nrows,rs,st);
#else
  return 0;
#endif
}

} // namespace

// Same contract as fast_try_convolve: 1 handled, 0 not applicable, <0 error.
int tmem_try_convolve(Plan *pl, void *const *f, uint32_t A, uint32_t B,
                      int mult, double scale, uint64_t nrows, uint64_t rs,
                      cudaStream_t st)
{
  {
    int rc=tryLongRows(pl,f,A,B,mult,scale,nrows,rs,st);
    if(rc != 0) return rc;
  }
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
  const int sms=sm_count();
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
