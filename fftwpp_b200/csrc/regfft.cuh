// regfft.cuh -- register-resident radix-8 FFT building blocks shared by the
// specialised power-of-two kernels (fast_kernels.cu, tma_kernels.cu): complex
// helpers, radix-8/4/2 butterflies, the RegFFT<LG> pass/exchange driver and
// the residue-twiddle lookups.  Everything has internal linkage.
#ifndef FFTWPP_REGFFT_CUH
#define FFTWPP_REGFFT_CUH

#include "gpu_internal.h"

#ifndef FFTWPP_TWMODE
#define FFTWPP_TWMODE 1
#endif

namespace fftwpp_gpu {
namespace {

__device__ __forceinline__ double2 fmul(double2 a, double2 b)
{
  return make_double2(fma(a.x,b.x,-a.y*b.y),fma(a.x,b.y,a.y*b.x));
}

__device__ __forceinline__ double2 fmulc(double2 a, double2 b) // a*conj(b)
{
  return make_double2(fma(a.x,b.x,a.y*b.y),fma(a.y,b.x,-a.x*b.y));
}

__device__ __forceinline__ double2 operator+(double2 a, double2 b)
{
  return make_double2(a.x+b.x,a.y+b.y);
}

__device__ __forceinline__ double2 operator-(double2 a, double2 b)
{
  return make_double2(a.x-b.x,a.y-b.y);
}

// multiply by SIGN*i
template<int SIGN>
__device__ __forceinline__ double2 rot(double2 a)
{
  return SIGN > 0 ? make_double2(-a.y,a.x) : make_double2(a.y,-a.x);
}

// y_v = sum_u a_u exp(SIGN 2 pi i u v/8), in place, natural order.
template<int SIGN>
__device__ __forceinline__ void bfly8(double2 (&a)[8])
{
  const double h=0.70710678118654752440;
  double2 b0=a[0]+a[4], b4=a[0]-a[4];
  double2 b1=a[1]+a[5], b5=a[1]-a[5];
  double2 b2=a[2]+a[6], b6=a[2]-a[6];
  double2 b3=a[3]+a[7], b7=a[3]-a[7];
  // even outputs: radix 4 on b0..b3
  double2 c0=b0+b2, c2=b0-b2;
  double2 c1=b1+b3, c3=rot<SIGN>(b1-b3);
  a[0]=c0+c1;
  a[4]=c0-c1;
  a[2]=c2+c3;
  a[6]=c2-c3;
  // odd outputs: radix 4 on b4, w b5, w^2 b6, w^3 b7 with w=exp(SIGN i pi/4);
  // the 1/sqrt(2) factors are folded into the final FMAs
  double2 r6=rot<SIGN>(b6);
  double2 d0=b4+r6, d2=b4-r6;
  double2 p5=b5+rot<SIGN>(b5);           // sqrt(2) w b5
  double2 p7=rot<SIGN>(b7)-b7;           // sqrt(2) w^3 b7
  double2 s1=p5+p7;
  double2 s3=rot<SIGN>(p5-p7);
  a[1]=make_double2(fma(h,s1.x,d0.x),fma(h,s1.y,d0.y));
  a[5]=make_double2(fma(-h,s1.x,d0.x),fma(-h,s1.y,d0.y));
  a[3]=make_double2(fma(h,s3.x,d2.x),fma(h,s3.y,d2.y));
  a[7]=make_double2(fma(-h,s3.x,d2.x),fma(-h,s3.y,d2.y));
}

template<int SIGN>
__device__ __forceinline__ void bfly4(double2& a0, double2& a1, double2& a2,
                                      double2& a3)
{
  double2 c0=a0+a2, c2=a0-a2;
  double2 c1=a1+a3, c3=rot<SIGN>(a1-a3);
  a0=c0+c1;
  a1=c2+c3;
  a2=c0-c1;
  a3=c2-c3;
}

__device__ __forceinline__ void bfly2(double2& a0, double2& a1)
{
  double2 t=a0-a1;
  a0=a0+a1;
  a1=t;
}

// Shared-memory addressing policies.
struct LaneLayout { // element p of lane `lane` at p*T+lane; CTA-wide barrier
  int T, lane;
  __device__ __forceinline__ int addr(int p) const {return p*T+lane;}
  __device__ __forceinline__ void sync() const {__syncthreads();}
};

struct RowLayout { // one padded buffer per row; barrier over the row's threads
  int base;        // row offset inside the exchange buffer (in double2)
  int barid;       // named barrier id (0: the row lives inside one warp)
  int nthreads;
  __device__ __forceinline__ int addr(int p) const {return base+p+(p >> 3);}
  __device__ __forceinline__ void sync() const {
    if(barid == 0) __syncwarp();
    else asm volatile("bar.sync %0, %1;" :: "r"(barid), "r"(nthreads) : "memory");
  }
};

// Register FFT of length N=2^LG over N/8 threads, NA arrays at a time (the
// arrays share twiddle loads and barriers).
template<int LG>
struct RegFFT {
  static const int N=1 << LG;
  static const int TPT=N/8;
  static const int NR8=LG/3;
  static const int REM=LG % 3;

  // offset of pass i in the tw8 table and number of table entries
  static __host__ __device__ constexpr int twOff(int i) {
    int off=0;
    for(int k=0; k < i; ++k) off += 7 << (LG-3*(k+1));
    return off;
  }
  static __host__ __device__ constexpr int twCount() {
    int off=0;
    for(int k=0; k < NR8; ++k)
      if(LG-3*(k+1) > 0) off += 7 << (LG-3*(k+1));
    return off;
  }
  template<bool SM>
  static __device__ __forceinline__ double2 twid(const double2 *tw, int i,
                                                 int u, int tau) {
    const int ls=LG-3*(i+1);
    const double2 *p=tw+twOff(i)+((u-1) << ls)+(tau & ((1 << ls)-1));
    return SM ? *p : __ldg(p);
  }

  // The seven twiddles w^1..w^7 of pass i.  Shared-memory bandwidth, not
  // FP64 issue, limits these kernels, so only w, w^2 and w^4 are read and the
  // other four are products of those (TWM 1, the strided passes), or w alone
  // is read (TWM 2, the fused convolutions); TWM 0 reads all seven.
  template<bool SM, int TWM>
  static __device__ __forceinline__ void twiddles(const double2 *tw, int i,
                                                  int tau, double2 (&w)[8]) {
    if(TWM == 0) {
#pragma unroll
      for(int u=1; u < 8; ++u) w[u]=twid<SM>(tw,i,u,tau);
    } else if(TWM == 1) {
      w[1]=twid<SM>(tw,i,1,tau);
      w[2]=twid<SM>(tw,i,2,tau);
      w[4]=twid<SM>(tw,i,4,tau);
      w[3]=fmul(w[1],w[2]);
      w[5]=fmul(w[1],w[4]);
      w[6]=fmul(w[2],w[4]);
      w[7]=fmul(w[3],w[4]);
    } else {
      w[1]=twid<SM>(tw,i,1,tau);
      w[2]=fmul(w[1],w[1]);
      w[3]=fmul(w[1],w[2]);
      w[4]=fmul(w[2],w[2]);
      w[5]=fmul(w[1],w[4]);
      w[6]=fmul(w[3],w[3]);
      w[7]=fmul(w[3],w[4]);
    }
  }

  // position of register t of thread tau in a pass whose legs are 2^ls apart
  static __device__ __forceinline__ int pos(int tau, int t, int ls) {
    return ((tau >> ls) << (ls+3))+(tau & ((1 << ls)-1))+(t << ls);
  }

  // Array a uses buffer buf+a*bufStride.  Default: barrier before the writes
  // (protects the previous readers) and after (publishes).  With PP the
  // exchanges alternate between two buffers ppStride apart, which makes the
  // leading barrier unnecessary (a buffer is rewritten only two exchanges
  // later, behind the intermediate exchange's barrier).
  template<int NA, class Lay, bool PP>
  static __device__ __forceinline__ void exchange(double2 (&x)[NA][8], int tau,
                                                  int lsFrom, int lsTo,
                                                  double2 *buf, int bufStride,
                                                  const Lay& lay, bool active,
                                                  int& pp, int ppStride) {
    if(PP && ppStride) {
      buf += pp*ppStride;
      pp ^= 1;
    } else
      lay.sync();
    if(active) {
#pragma unroll
      for(int a=0; a < NA; ++a)
#pragma unroll
        for(int t=0; t < 8; ++t)
          buf[a*bufStride+lay.addr(pos(tau,t,lsFrom))]=x[a][t];
    }
    lay.sync();
    if(active) {
#pragma unroll
      for(int a=0; a < NA; ++a)
#pragma unroll
        for(int t=0; t < 8; ++t)
          x[a][t]=buf[a*bufStride+lay.addr(pos(tau,t,lsTo))];
    }
  }

  // in: x[a][t]=W_a[tau+TPT*t]; out: x[a][e]=FFT at scrambled position 8*tau+e
  template<int NA, class Lay, bool SM=false, bool PP=false,
           int TWM=FFTWPP_TWMODE>
  static __device__ __forceinline__ void forward(double2 (&x)[NA][8], int tau,
                                                 const double2 *__restrict__ tw,
                                                 double2 *buf, int bufStride,
                                                 const Lay& lay, bool active,
                                                 int *pp=NULL,
                                                 int ppStride=0) {
    int dummy=0;
    int& ppr=PP ? *pp : dummy;
#pragma unroll
    for(int i=0; i < NR8; ++i) {
      const int ls=LG-3*(i+1);
#pragma unroll
      for(int a=0; a < NA; ++a) bfly8<1>(x[a]);
      if(ls > 0) {
        double2 w[8];
        twiddles<SM,TWM>(tw,i,tau,w);
#pragma unroll
        for(int u=1; u < 8; ++u)
#pragma unroll
          for(int a=0; a < NA; ++a) x[a][u]=fmul(x[a][u],w[u]);
      }
      const int lsNext=(i+1 < NR8) ? LG-3*(i+2) : 0;
      if(i+1 < NR8 || REM > 0)
        exchange<NA,Lay,PP>(x,tau,ls,lsNext,buf,bufStride,lay,active,ppr,
                            ppStride);
    }
#pragma unroll
    for(int a=0; a < NA; ++a) {
      if(REM == 2) {
        bfly4<1>(x[a][0],x[a][1],x[a][2],x[a][3]);
        bfly4<1>(x[a][4],x[a][5],x[a][6],x[a][7]);
      } else if(REM == 1) {
        bfly2(x[a][0],x[a][1]);
        bfly2(x[a][2],x[a][3]);
        bfly2(x[a][4],x[a][5]);
        bfly2(x[a][6],x[a][7]);
      }
    }
  }

  // exact adjoint of forward(): in scrambled positions, out x[t]=w[tau+TPT*t]
  template<int NA, class Lay, bool SM=false, bool PP=false,
           int TWM=FFTWPP_TWMODE>
  static __device__ __forceinline__ void adjoint(double2 (&x)[NA][8], int tau,
                                                 const double2 *__restrict__ tw,
                                                 double2 *buf, int bufStride,
                                                 const Lay& lay, bool active,
                                                 int *pp=NULL,
                                                 int ppStride=0) {
    int dummy=0;
    int& ppr=PP ? *pp : dummy;
#pragma unroll
    for(int a=0; a < NA; ++a) {
      if(REM == 2) {
        bfly4<-1>(x[a][0],x[a][1],x[a][2],x[a][3]);
        bfly4<-1>(x[a][4],x[a][5],x[a][6],x[a][7]);
      } else if(REM == 1) {
        bfly2(x[a][0],x[a][1]);
        bfly2(x[a][2],x[a][3]);
        bfly2(x[a][4],x[a][5]);
        bfly2(x[a][6],x[a][7]);
      }
    }
#pragma unroll
    for(int i=NR8-1; i >= 0; --i) {
      const int ls=LG-3*(i+1);
      const int lsPrev=(i+1 < NR8) ? LG-3*(i+2) : 0;
      if(i+1 < NR8 || REM > 0)
        exchange<NA,Lay,PP>(x,tau,lsPrev,ls,buf,bufStride,lay,active,ppr,
                            ppStride);
      if(ls > 0) {
        double2 w[8];
        twiddles<SM,TWM>(tw,i,tau,w);
#pragma unroll
        for(int u=1; u < 8; ++u)
#pragma unroll
          for(int a=0; a < NA; ++a) x[a][u]=fmulc(x[a][u],w[u]);
      }
#pragma unroll
      for(int a=0; a < NA; ++a) bfly8<-1>(x[a]);
    }
  }

  // transformed index held at scrambled position p (mixed-radix digit reversal)
  static __device__ __forceinline__ int rev(int p) {
    int rem=p, l=0, shift=0, lg=LG;
#pragma unroll
    for(int i=0; i < NR8; ++i) {
      lg -= 3;
      l += (rem >> lg) << shift;
      rem &= (1 << lg)-1;
      shift += 3;
    }
    if(REM > 0) l += rem << shift;
    return l;
  }
};

__device__ __forceinline__ double2 zeta(const PlanDev& P, long long e)
{
  if(P.zshift < 0) return __ldg(P.z1+e);
  long long hi=e >> P.zshift;
  long long lo=e & ((1ll << P.zshift)-1);
  return fmul(__ldg(P.z1+hi),__ldg(P.z2+lo));
}

// (k0*j) mod N without 64-bit division on the common paths
__device__ __forceinline__ long long modN(const PlanDev& P, long long k0,
                                          int j)
{
  if(P.nmask) return (long long) (((unsigned) k0*(unsigned) j) & P.nmask);
  if(P.small32) {
    unsigned N=(unsigned) P.N;
    unsigned e=((unsigned) k0*(unsigned) (j < 0 ? -j : j)) % N;
    if(j < 0 && e) e=N-e;
    return (long long) e;
  }
  long long e=(k0*j) % P.N;
  return e < 0 ? e+P.N : e;
}

// outer twiddle zeta_Nbig^e of a two-stage transform (parent plan's tables)
__device__ __forceinline__ double2 ozeta(const PlanDev& P, long long e)
{
  if(P.ozshift < 0) return __ldg(P.oz1+e);
  long long hi=e >> P.ozshift;
  long long lo=e & ((1ll << P.ozshift)-1);
  return fmul(__ldg(P.oz1+hi),__ldg(P.oz2+lo));
}


// Register FFT with the radix-8 twiddle base of every pass held in registers
// (persistent CTAs: loaded once); w^2..w^7 are products.
template<int LG>
struct WFFT {
  typedef RegFFT<LG> F;
  static const int NR8=F::NR8;

  static __device__ __forceinline__ void loadW(const double2 *tw8, int tau,
                                               double2 (&w1)[NR8 > 0 ? NR8 : 1])
  {
#pragma unroll
    for(int i=0; i < NR8; ++i) {
      const int ls=LG-3*(i+1);
      w1[i]=ls > 0 ? __ldg(tw8+F::twOff(i)+(tau & ((1 << ls)-1))) :
        make_double2(1.0,0.0);
    }
  }

  static __device__ __forceinline__ void powers(double2 w1, double2 (&w)[8])
  {
    w[1]=w1;
    w[2]=fmul(w[1],w[1]);
    w[3]=fmul(w[1],w[2]);
    w[4]=fmul(w[2],w[2]);
    w[5]=fmul(w[1],w[4]);
    w[6]=fmul(w[3],w[3]);
    w[7]=fmul(w[3],w[4]);
  }

  template<class Lay>
  static __device__ __forceinline__ void exchange(double2 (&x)[8], int tau,
                                                  int lsFrom, int lsTo,
                                                  double2 *buf, const Lay& lay)
  {
    lay.sync();
#pragma unroll
    for(int t=0; t < 8; ++t) buf[lay.addr(F::pos(tau,t,lsFrom))]=x[t];
    lay.sync();
#pragma unroll
    for(int t=0; t < 8; ++t) x[t]=buf[lay.addr(F::pos(tau,t,lsTo))];
  }

  // in: x[t]=W[tau+TPT*t]; out: x[e] at scrambled position 8*tau+e
  template<class Lay>
  static __device__ __forceinline__ void forward(double2 (&x)[8], int tau,
                                                 const double2 (&w1)[NR8 > 0 ? NR8 : 1],
                                                 double2 *buf, const Lay& lay)
  {
#pragma unroll
    for(int i=0; i < NR8; ++i) {
      const int ls=LG-3*(i+1);
      bfly8<1>(x);
      if(ls > 0) {
        double2 w[8];
        powers(w1[i],w);
#pragma unroll
        for(int u=1; u < 8; ++u) x[u]=fmul(x[u],w[u]);
      }
      const int lsNext=(i+1 < NR8) ? LG-3*(i+2) : 0;
      if(i+1 < NR8 || F::REM > 0) exchange(x,tau,ls,lsNext,buf,lay);
    }
    if(F::REM == 2) {
      bfly4<1>(x[0],x[1],x[2],x[3]);
      bfly4<1>(x[4],x[5],x[6],x[7]);
    } else if(F::REM == 1) {
      bfly2(x[0],x[1]);
      bfly2(x[2],x[3]);
      bfly2(x[4],x[5]);
      bfly2(x[6],x[7]);
    }
  }

  // exact adjoint of forward()
  template<class Lay>
  static __device__ __forceinline__ void adjoint(double2 (&x)[8], int tau,
                                                 const double2 (&w1)[NR8 > 0 ? NR8 : 1],
                                                 double2 *buf, const Lay& lay)
  {
    if(F::REM == 2) {
      bfly4<-1>(x[0],x[1],x[2],x[3]);
      bfly4<-1>(x[4],x[5],x[6],x[7]);
    } else if(F::REM == 1) {
      bfly2(x[0],x[1]);
      bfly2(x[2],x[3]);
      bfly2(x[4],x[5]);
      bfly2(x[6],x[7]);
    }
#pragma unroll
    for(int i=NR8-1; i >= 0; --i) {
      const int ls=LG-3*(i+1);
      const int lsPrev=(i+1 < NR8) ? LG-3*(i+2) : 0;
      if(i+1 < NR8 || F::REM > 0) exchange(x,tau,lsPrev,ls,buf,lay);
      if(ls > 0) {
        double2 w[8];
        powers(w1[i],w);
#pragma unroll
        for(int u=1; u < 8; ++u) x[u]=fmulc(x[u],w[u]);
      }
      bfly8<-1>(x);
    }
  }
};

} // namespace
} // namespace fftwpp_gpu

#endif
