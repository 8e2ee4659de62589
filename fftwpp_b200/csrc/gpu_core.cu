// gpu_core.cu -- plans, tables and the GENERIC residue-pass kernels behind the
// thin C ABI (include/fftwpp_gpu.h).  sm_100a only.
//
// One mathematical formulation covers every reference routine family
// (convolve.cc:771-7483): each forward(r) call of fftPad / fftPadCentered /
// fftPadHermitian / fftPadReal is a list of "sub-blocks"
//     W[s]   = sum_{j = s (mod m')} zeta_N^{k0 j} g(j)          (prologue)
//     out[l] = sum_s zeta_{m'}^{l s} W[s]                        (FFT_{m'})
// and each backward(r) call is its adjoint.  The kernels below are generic in
// m' (any product of primes <= 64, mixed radix), p (any number of folded
// terms), the four input kinds and arbitrary stride/count, and hold the padded
// data in shared memory only.  Power-of-two hot shapes are dispatched to the
// specialised register kernels in fast_kernels.cu.
//
// Shared-memory tile layout: T independent transforms ("lanes": columns of a
// strided "Many" pass, or rows of a contiguous batch) are interleaved lane-
// fastest, element (s,t) at s*T+t, so every butterfly stage is bank-conflict
// free and global accesses of strided passes are whole 128-byte lines.
// The forward FFT is an in-place decimation-in-frequency transform that leaves
// its output in mixed-radix digit-reversed order; the backward transform is
// its exact adjoint and consumes that order, so the fused convolution never
// reorders anything (the multiplier is pointwise).

#include "gpu_internal.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

namespace fftwpp_gpu {

thread_local char g_err[512]="";
std::atomic<uint64_t> g_launches(0);

void set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap,fmt);
  vsnprintf(g_err,sizeof(g_err),fmt,ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
  set_error("%s: %s",what,cudaGetErrorString(e));
  if(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
    return FFTWPP_GPU_ENODEVICE;
  if(e == cudaErrorMemoryAllocation)
    return FFTWPP_GPU_ENOMEM;
  return FFTWPP_GPU_ECUDA;
}

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
  return make_double2(fma(a.x,b.x,-a.y*b.y),fma(a.x,b.y,a.y*b.x));
}

// a*conj(b)
__device__ __forceinline__ double2 cmulc(double2 a, double2 b)
{
  return make_double2(fma(a.x,b.x,a.y*b.y),fma(a.y,b.x,-a.x*b.y));
}

__device__ __forceinline__ double2 cadd(double2 a, double2 b)
{
  return make_double2(a.x+b.x,a.y+b.y);
}

__device__ __forceinline__ double2 csub(double2 a, double2 b)
{
  return make_double2(a.x-b.x,a.y-b.y);
}

__device__ __forceinline__ double2 zetaN(const PlanDev& P, long long e)
{
  if(P.zshift < 0) return P.z1[e];
  long long hi=e >> P.zshift;
  long long lo=e & ((1ll << P.zshift)-1);
  return cmul(P.z1[hi],P.z2[lo]);
}

__device__ __forceinline__ int digitrev(const FftTab& tab, int pos)
{
  int len=tab.n;
  int l=0;
  int mult=1;
  int rem=pos;
  for(int i=0; i < tab.nrad; ++i) {
    int r=tab.rad[i];
    len /= r;
    int t=rem/len;
    rem -= t*len;
    l += t*mult;
    mult *= r;
  }
  return l;
}

// One in-place decimation-in-frequency stage (ADJ=false, exponent sign +) or
// its adjoint (ADJ=true) on `narr` arrays of T interleaved lanes.
template<bool ADJ>
__device__ void fft_stage(double2 *w, size_t arrstride, int narr, int T,
                          const FftTab& tab, int len, int r)
{
  const int n=tab.n;
  const int sub=len/r;
  const int nbf=n/r;
  const int tstep=n/len;
  const int rstep=n/r;
  const int total=nbf*T*narr;
  for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
    int t=idx % T;
    int q=idx/T;
    int bf=q % nbf;
    int arr=q/nbf;
    int blk=bf/sub;
    int j=bf-blk*sub;
    double2 *p=w+arr*arrstride+(size_t) (blk*len+j)*T+t;
    size_t leg=(size_t) sub*T;
    if(r == 2) {
      double2 a0=p[0], a1=p[leg];
      if(!ADJ) {
        double2 y1=csub(a0,a1);
        if(j) y1=cmul(y1,tab.omega[j*tstep]);
        p[0]=cadd(a0,a1);
        p[leg]=y1;
      } else {
        if(j) a1=cmulc(a1,tab.omega[j*tstep]);
        p[0]=cadd(a0,a1);
        p[leg]=csub(a0,a1);
      }
    } else if(r == 4) {
      double2 a0=p[0], a1=p[leg], a2=p[2*leg], a3=p[3*leg];
      if(!ADJ) {
        double2 s02=cadd(a0,a2), d02=csub(a0,a2);
        double2 s13=cadd(a1,a3), d13=csub(a1,a3);
        double2 id13=make_double2(-d13.y,d13.x); // +i*d13
        double2 y0=cadd(s02,s13);
        double2 y1=cadd(d02,id13);
        double2 y2=csub(s02,s13);
        double2 y3=csub(d02,id13);
        if(j) {
          y1=cmul(y1,tab.omega[j*tstep]);
          y2=cmul(y2,tab.omega[2*j*tstep]);
          y3=cmul(y3,tab.omega[3*j*tstep]);
        }
        p[0]=y0; p[leg]=y1; p[2*leg]=y2; p[3*leg]=y3;
      } else {
        if(j) {
          a1=cmulc(a1,tab.omega[j*tstep]);
          a2=cmulc(a2,tab.omega[2*j*tstep]);
          a3=cmulc(a3,tab.omega[3*j*tstep]);
        }
        double2 s02=cadd(a0,a2), d02=csub(a0,a2);
        double2 s13=cadd(a1,a3), d13=csub(a1,a3);
        double2 id13=make_double2(d13.y,-d13.x); // -i*d13
        p[0]=cadd(s02,s13);
        p[leg]=cadd(d02,id13);
        p[2*leg]=csub(s02,s13);
        p[3*leg]=csub(d02,id13);
      }
    } else {
      double2 a[MAXPRIME];
      double2 y[MAXPRIME];
      for(int u=0; u < r; ++u) {
        double2 v=p[u*leg];
        if(ADJ && j && u) v=cmulc(v,tab.omega[(size_t) j*u*tstep]);
        a[u]=v;
      }
      for(int v=0; v < r; ++v) {
        double2 sum=a[0];
        for(int u=1; u < r; ++u) {
          double2 wr=tab.omega[(size_t) ((u*v) % r)*rstep];
          sum=cadd(sum,ADJ ? cmulc(a[u],wr) : cmul(a[u],wr));
        }
        if(!ADJ && j && v) sum=cmul(sum,tab.omega[(size_t) j*v*tstep]);
        y[v]=sum;
      }
      for(int v=0; v < r; ++v)
        p[v*leg]=y[v];
    }
  }
}

__device__ void fft_forward(double2 *w, size_t arrstride, int narr, int T,
                            const FftTab& tab)
{
  int len=tab.n;
  for(int i=0; i < tab.nrad; ++i) {
    int r=tab.rad[i];
    fft_stage<false>(w,arrstride,narr,T,tab,len,r);
    __syncthreads();
    len /= r;
  }
}

__device__ void fft_adjoint(double2 *w, size_t arrstride, int narr, int T,
                            const FftTab& tab)
{
  int len=1;
  for(int i=tab.nrad-1; i >= 0; --i) {
    int r=tab.rad[i];
    len *= r;
    fft_stage<true>(w,arrstride,narr,T,tab,len,r);
    __syncthreads();
  }
}

// Logical input sample j of lane t from the staged tile.
template<int KIND>
__device__ __forceinline__ double2 ginput(const void *in, int j, int jmin,
                                          int T, int t)
{
  if(KIND == FFTWPP_KIND_REAL)
    return make_double2(((const double *) in)[(size_t) j*T+t],0.0);
  const double2 *c=(const double2 *) in;
  if(KIND == FFTWPP_KIND_HERMITIAN) {
    if(j >= 0) return c[(size_t) j*T+t];
    double2 v=c[(size_t) (-j)*T+t];
    return make_double2(v.x,-v.y);
  }
  return c[(size_t) (j-jmin)*T+t];
}

// W[s]=sum_{j=s mod mlen} zeta^{k0 j} g(j) for all lanes of one array.
template<int KIND>
__device__ void build_W(const PlanDev& P, const SubBlockDev& sb,
                        const void *in, double2 *w, int T)
{
  const int mlen=sb.mlen;
  const int jmin=P.jmin, jmax=P.jmax;
  const long long N=P.N;
  const long long k0=sb.k0;
  const int total=mlen*T;
  for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
    int t=idx % T;
    int s=idx/T;
    int d=(s-jmin) % mlen;
    double2 acc=make_double2(0.0,0.0);
    for(int j=jmin+d; j < jmax; j += mlen) {
      double2 v=ginput<KIND>(in,j,jmin,T,t);
      if(k0 != 0) {
        long long e=(k0*j) % N;
        if(e < 0) e += N;
        v=cmul(v,zetaN(P,e));
      }
      acc=cadd(acc,v);
    }
    w[idx]=acc;
  }
}

template<int KIND>
struct InWord {typedef double2 type;};
template<>
struct InWord<FFTWPP_KIND_REAL> {typedef double type;};

// Stage one input array tile: in[j*T+t] = f[row*rs + S*j + col].
template<int KIND>
__device__ void stage_input(const PlanDev& P, const void *f, void *in,
                            long long row0, long long nrows, long long rs,
                            int col0, int TR, int TC)
{
  typedef typename InWord<KIND>::type word;
  const word *g=(const word *) f;
  word *dst=(word *) in;
  const int T=TR*TC;
  const int Lin=P.Lin;
  const int total=Lin*T;
  if(TC > 1) {
    for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
      int t=idx % T;
      int j=idx/T;
      int rr=t/TC, cc=t-rr*TC;
      long long row=row0+rr;
      int col=col0+cc;
      word v=word();
      if(row < nrows && col < P.C)
        v=g[row*rs+P.S*j+col];
      dst[idx]=v;
    }
  } else {
    // rows mode: iterate j fastest so global reads of a row are contiguous
    for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
      int j=idx % Lin;
      int t=idx/Lin;
      long long row=row0+t;
      word v=word();
      if(row < nrows)
        v=g[row*rs+P.S*j+col0];
      dst[(size_t) j*T+t]=v;
    }
  }
}

template<int KIND>
__global__ void __launch_bounds__(NTHREADS)
forward_kernel(PlanDev P, const SubBlockDev *sbs, int nsb, int layout,
               const void *f, void *F, long long nrows, long long frs,
               long long Frs, int TR, int TC, int ntc, size_t inbytes)
{
  extern __shared__ __align__(16) unsigned char smem[];
  void *in=smem;
  double2 *w=(double2 *) (smem+inbytes);
  const int T=TR*TC;
  const long long row0=(long long) (blockIdx.x/ntc)*TR;
  const int col0=(blockIdx.x % ntc)*TC;

  stage_input<KIND>(P,f,in,row0,nrows,frs,col0,TR,TC);
  __syncthreads();

  for(int isb=0; isb < nsb; ++isb) {
    const SubBlockDev sb=sbs[isb];
    const FftTab& tab=P.tab[sb.tab];
    build_W<KIND>(P,sb,in,w,T);
    __syncthreads();
    fft_forward(w,0,1,T,tab);
    const long long off=layout ? sb.off_all : sb.off_call;
    const int total=sb.mlen*T;
    for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
      int t=idx % T;
      int pos=idx/T;
      int l=digitrev(tab,pos);
      if(l >= (int) sb.nout) continue;
      int rr=t/TC, cc=t-rr*TC;
      long long row=row0+rr;
      int col=col0+cc;
      if(row >= nrows || col >= P.C) continue;
      double2 v=w[idx];
      long long a=row*Frs+off+P.S*l+col;
      if(KIND == FFTWPP_KIND_HERMITIAN)
        ((double *) F)[a]=v.x;
      else {
        if(sb.flags & FFTWPP_SB_CONJ_OUT) v.y=-v.y;
        ((double2 *) F)[a]=v;
      }
    }
    __syncthreads();
  }
}

// Load one transformed sub-block from global memory into digit-reversed
// shared-memory order, rebuilding the implied half of r2c blocks.
template<int KIND>
__device__ void load_F(const PlanDev& P, const SubBlockDev& sb,
                       const FftTab& tab, const void *F, double2 *w,
                       long long off, long long row0, long long nrows,
                       long long Frs, int col0, int TR, int TC)
{
  const int T=TR*TC;
  const int total=sb.mlen*T;
  for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
    int t=idx % T;
    int pos=idx/T;
    int l=digitrev(tab,pos);
    int rr=t/TC, cc=t-rr*TC;
    long long row=row0+rr;
    int col=col0+cc;
    double2 v=make_double2(0.0,0.0);
    if(row < nrows && col < P.C) {
      long long base=row*Frs+off+col;
      if(KIND == FFTWPP_KIND_HERMITIAN)
        v.x=((const double *) F)[base+P.S*l];
      else if(sb.flags & FFTWPP_SB_CONJ_OUT) {
        // stored block holds conj(G[l]), l < nout; G[mlen-l]=conj(G[l])
        if(l < (int) sb.nout) {
          v=((const double2 *) F)[base+P.S*l];
          v.y=-v.y;
        } else
          v=((const double2 *) F)[base+P.S*(sb.mlen-l)];
      } else
        v=((const double2 *) F)[base+P.S*l];
    }
    w[idx]=v;
  }
}

// acc[j] += conj(zeta^{k0 j}) w[j mod mlen] over the stored input range.
template<int KIND>
__device__ void accumulate(const PlanDev& P, const SubBlockDev& sb,
                           const double2 *w, void *acc, int T)
{
  const int Lin=P.Lin;
  const int mlen=sb.mlen;
  const long long N=P.N;
  const long long k0=sb.k0;
  const int total=Lin*T;
  const int shift=(KIND == FFTWPP_KIND_CENTERED) ? P.jmin : 0;
  for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
    int t=idx % T;
    int jj=idx/T;
    int j=jj+shift;
    int s=j % mlen;
    if(s < 0) s += mlen;
    double2 v=w[(size_t) s*T+t];
    if(k0 != 0) {
      long long e=(k0*j) % N;
      if(e < 0) e += N;
      v=cmulc(v,zetaN(P,e));
    }
    if(KIND == FFTWPP_KIND_REAL) {
      double *a=(double *) acc;
      a[idx] += (sb.flags & (FFTWPP_SB_CONJ_OUT | FFTWPP_SB_SELFCONJ))
          ? v.x : 2.0*v.x;
    } else {
      double2 *a=(double2 *) acc;
      a[idx]=cadd(a[idx],v);
    }
  }
}

template<int KIND>
__device__ void store_acc(const PlanDev& P, const void *acc, void *f,
                          double scale, long long row0, long long nrows,
                          long long rs, int col0, int TR, int TC)
{
  typedef typename InWord<KIND>::type word;
  const word *src=(const word *) acc;
  word *g=(word *) f;
  const int T=TR*TC;
  const int Lin=P.Lin;
  const int total=Lin*T;
  if(TC > 1) {
    for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
      int t=idx % T;
      int j=idx/T;
      int rr=t/TC, cc=t-rr*TC;
      long long row=row0+rr;
      int col=col0+cc;
      if(row < nrows && col < P.C)
        g[row*rs+P.S*j+col]=wscale(src[idx],scale);
    }
  } else {
    for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
      int j=idx % Lin;
      int t=idx/Lin;
      long long row=row0+t;
      if(row < nrows)
        g[row*rs+P.S*j+col0]=wscale(src[(size_t) j*T+t],scale);
    }
  }
}

template<int KIND>
__global__ void __launch_bounds__(NTHREADS)
backward_kernel(PlanDev P, const SubBlockDev *sbs, int nsb, int layout,
                const void *F, void *f, int accum, double scale,
                long long nrows, long long Frs, long long frs, int TR, int TC,
                int ntc, size_t accbytes)
{
  typedef typename InWord<KIND>::type word;
  extern __shared__ __align__(16) unsigned char smem[];
  void *acc=smem;
  double2 *w=(double2 *) (smem+accbytes);
  const int T=TR*TC;
  const long long row0=(long long) (blockIdx.x/ntc)*TR;
  const int col0=(blockIdx.x % ntc)*TC;

  if(accum)
    stage_input<KIND>(P,f,acc,row0,nrows,frs,col0,TR,TC);
  else {
    word *a=(word *) acc;
    for(int idx=threadIdx.x; idx < P.Lin*T; idx += blockDim.x)
      a[idx]=word();
  }
  __syncthreads();

  for(int isb=0; isb < nsb; ++isb) {
    const SubBlockDev sb=sbs[isb];
    const FftTab& tab=P.tab[sb.tab];
    const long long off=layout ? sb.off_all : sb.off_call;
    load_F<KIND>(P,sb,tab,F,w,off,row0,nrows,Frs,col0,TR,TC);
    __syncthreads();
    fft_adjoint(w,0,1,T,tab);
    accumulate<KIND>(P,sb,w,acc,T);
    __syncthreads();
  }
  store_acc<KIND>(P,acc,f,scale,row0,nrows,frs,col0,TR,TC);
}

// Fused 1-D convolution over TR rows per CTA (C == 1).
template<int KIND>
__global__ void __launch_bounds__(NTHREADS)
convolve_kernel(PlanDev P, const SubBlockDev *sbs, int nsb, ConvPtrs ptrs,
                int A, int B, int mult, double scale, long long nrows,
                long long rs, int TR, size_t inbytes, size_t wstride)
{
  typedef typename InWord<KIND>::type word;
  extern __shared__ __align__(16) unsigned char smem[];
  const int T=TR;
  const long long row0=(long long) blockIdx.x*TR;
  unsigned char *in=smem;                       // A tiles of inbytes
  unsigned char *acc=smem+(size_t) A*inbytes;   // B tiles of inbytes
  double2 *w=(double2 *) (acc+(size_t) B*inbytes); // max(A,B) arrays of wstride

  for(int a=0; a < A; ++a)
    stage_input<KIND>(P,ptrs.p[a],in+(size_t) a*inbytes,row0,nrows,rs,0,TR,1);
  for(int b=0; b < B; ++b) {
    word *ab=(word *) (acc+(size_t) b*inbytes);
    for(int idx=threadIdx.x; idx < P.Lin*T; idx += blockDim.x)
      ab[idx]=word();
  }
  __syncthreads();

  for(int isb=0; isb < nsb; ++isb) {
    const SubBlockDev sb=sbs[isb];
    const FftTab& tab=P.tab[sb.tab];
    for(int a=0; a < A; ++a)
      build_W<KIND>(P,sb,in+(size_t) a*inbytes,w+(size_t) a*wstride,T);
    __syncthreads();
    fft_forward(w,wstride,A,T,tab);
    const int total=sb.mlen*T;
    if(mult == FFTWPP_MULT_BINARY) {
      for(int idx=threadIdx.x; idx < total; idx += blockDim.x)
        w[idx]=cmul(w[idx],w[wstride+idx]);
    } else if(mult == FFTWPP_MULT_CORRELATION) {
      for(int idx=threadIdx.x; idx < total; idx += blockDim.x)
        w[idx]=cmulc(w[idx],w[wstride+idx]);
    } else if(mult == FFTWPP_MULT_REALBINARY) {
      // real transforms: the multiplier sees Re only (convolve.cc:60-83)
      for(int idx=threadIdx.x; idx < total; idx += blockDim.x)
        w[idx]=make_double2(w[idx].x*w[wstride+idx].x,0.0);
    } else if(KIND == FFTWPP_KIND_HERMITIAN) {
      for(int b=0; b < B; ++b)
        for(int idx=threadIdx.x; idx < total; idx += blockDim.x)
          w[(size_t) b*wstride+idx].y=0.0;
    }
    __syncthreads();
    fft_adjoint(w,wstride,B,T,tab);
    for(int b=0; b < B; ++b)
      accumulate<KIND>(P,sb,w+(size_t) b*wstride,acc+(size_t) b*inbytes,T);
    __syncthreads();
  }
  for(int b=0; b < B; ++b)
    store_acc<KIND>(P,acc+(size_t) b*inbytes,ptrs.p[b],scale,row0,nrows,rs,0,
                    TR,1);
}

__global__ void scale_kernel(double *x, double scale, unsigned long long n0,
                             unsigned long long n1, unsigned long long n2,
                             unsigned long long s0, unsigned long long s1)
{
  unsigned long long total=n0*n1*n2;
  for(unsigned long long idx=blockIdx.x*(unsigned long long) blockDim.x+threadIdx.x;
      idx < total; idx += (unsigned long long) gridDim.x*blockDim.x) {
    unsigned long long k=idx % n2;
    unsigned long long q=idx/n2;
    unsigned long long j=q % n1;
    unsigned long long i=q/n1;
    x[i*s0+j*s1+k] *= scale;
  }
}

__global__ void copy3_kernel(double2 *dst, const double2 *src,
                             unsigned long long n0, unsigned long long n1,
                             unsigned long long n2, unsigned long long d0,
                             unsigned long long d1, unsigned long long s0,
                             unsigned long long s1)
{
  unsigned long long total=n0*n1*n2;
  for(unsigned long long idx=blockIdx.x*(unsigned long long) blockDim.x+threadIdx.x;
      idx < total; idx += (unsigned long long) gridDim.x*blockDim.x) {
    unsigned long long k=idx % n2;
    unsigned long long q=idx/n2;
    unsigned long long j=q % n1;
    unsigned long long i=q/n1;
    dst[i*d0+j*d1+k]=src[i*s0+j*s1+k];
  }
}

// ---------------------------------------------------------------------------
// host side: plan construction
// ---------------------------------------------------------------------------

static bool factorize(int n, FftTab& tab)
{
  tab.n=n;
  tab.nrad=0;
  int r=n;
  while(r % 4 == 0) {
    if(tab.nrad >= MAXRAD) return false;
    tab.rad[tab.nrad++]=4;
    r /= 4;
  }
  while(r % 2 == 0) {
    if(tab.nrad >= MAXRAD) return false;
    tab.rad[tab.nrad++]=2;
    r /= 2;
  }
  for(int f=3; f <= r; f += 2)
    while(r % f == 0) {
      if(f > MAXPRIME || tab.nrad >= MAXRAD) return false;
      tab.rad[tab.nrad++]=f;
      r /= f;
    }
  return true;
}

static int upload_roots(size_t count, size_t period, const double2 **out,
                        std::vector<void *>& owned, size_t mult=1)
{
  // table[k]=exp(2 pi i k*mult/period), k < count, evaluated in long double
  // with exact octant reduction of the integer phase.
  std::vector<double2> h(count);
  const long double twopi=6.283185307179586476925286766559005768L;
  for(size_t k=0; k < count; ++k) {
    unsigned long long ph=((unsigned long long) k*mult) % period;
    long double a=twopi*(long double) ph/(long double) period;
    h[k].x=(double) cosl(a);
    h[k].y=(double) sinl(a);
    // exact values on the axes
    if(4*ph == period) {h[k].x=0.0; h[k].y=1.0;}
    else if(2*ph == period) {h[k].x=-1.0; h[k].y=0.0;}
    else if(4*ph == 3*period) {h[k].x=0.0; h[k].y=-1.0;}
  }
  void *d=NULL;
  cudaError_t e=cudaMalloc(&d,count*sizeof(double2));
  if(e != cudaSuccess) return cuda_fail(e,"cudaMalloc(roots)");
  e=cudaMemcpy(d,h.data(),count*sizeof(double2),cudaMemcpyHostToDevice);
  if(e != cudaSuccess) {cudaFree(d); return cuda_fail(e,"cudaMemcpy(roots)");}
  owned.push_back(d);
  *out=(const double2 *) d;
  return FFTWPP_GPU_OK;
}

// Twiddles of the register radix-8 passes (see FftTab): pass i (legs 2^ls_i
// apart, ls_i=lg-3(i+1) > 0) needs omega^{(j*u) 8^i}, j < 2^ls_i, u=1..7,
// stored at off_i+(u-1)*2^ls_i+j with off_i=7*sum_{k<i} 2^ls_k.
static int upload_tw8(unsigned n, const double2 **out,
                      std::vector<void *>& owned)
{
  *out=NULL;
  if(n < 16 || (n & (n-1))) return FFTWPP_GPU_OK;
  int lg=0;
  while((1u << lg) < n) ++lg;
  int nr8=lg/3;
  std::vector<double2> h;
  const long double twopi=6.283185307179586476925286766559005768L;
  for(int i=0; i < nr8; ++i) {
    int ls=lg-3*(i+1);
    if(ls <= 0) break;
    for(int u=1; u < 8; ++u)
      for(unsigned j=0; j < (1u << ls); ++j) {
        unsigned long long ph=(((unsigned long long) j*u) << (3*i)) % n;
        long double a=twopi*(long double) ph/(long double) n;
        double2 v;
        v.x=(double) cosl(a);
        v.y=(double) sinl(a);
        if(ph == 0) {v.x=1.0; v.y=0.0;}
        else if(4*ph == n) {v.x=0.0; v.y=1.0;}
        else if(2*ph == n) {v.x=-1.0; v.y=0.0;}
        else if(4*ph == 3ull*n) {v.x=0.0; v.y=-1.0;}
        h.push_back(v);
      }
  }
  if(h.empty()) return FFTWPP_GPU_OK;
  void *d=NULL;
  cudaError_t e=cudaMalloc(&d,h.size()*sizeof(double2));
  if(e != cudaSuccess) return cuda_fail(e,"cudaMalloc(tw8)");
  e=cudaMemcpy(d,h.data(),h.size()*sizeof(double2),cudaMemcpyHostToDevice);
  if(e != cudaSuccess) {cudaFree(d); return cuda_fail(e,"cudaMemcpy(tw8)");}
  owned.push_back(d);
  *out=(const double2 *) d;
  return FFTWPP_GPU_OK;
}

Plan::~Plan()
{
  for(size_t i=0; i < owned.size(); ++i)
    cudaFree(owned[i]);
}

int plan_build(const fftwpp_gpu_pad_desc *d, Plan **out)
{
  if(!d || !out || d->nsub == 0 || !d->sub || d->m == 0 || d->N == 0 ||
     d->C == 0 || d->S < d->C || d->kind < 0 || d->kind > 3) {
    set_error("plan_create: invalid descriptor");
    return FFTWPP_GPU_EINVAL;
  }
  if(d->N > (1ull << 40) || d->Lin > (1ull << 30)) {
    set_error("plan_create: size too large");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  int ndev=0;
  cudaError_t ce=cudaGetDeviceCount(&ndev);
  if(ce != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (the GPU path has no CPU fallback)");
    return FFTWPP_GPU_ENODEVICE;
  }

  Plan *pl=new Plan;
  pl->desc=*d;
  pl->desc.sub=NULL;
  PlanDev& P=pl->dev;
  memset(&P,0,sizeof(P));
  P.kind=d->kind;
  P.Lin=(int) d->Lin;
  P.N=(long long) d->N;
  P.S=(long long) d->S;
  P.C=(int) d->C;
  switch(d->kind) {
    case FFTWPP_KIND_COMPLEX:
    case FFTWPP_KIND_REAL:
      P.jmin=0; P.jmax=(int) d->L; break;
    case FFTWPP_KIND_CENTERED:
      P.jmin=-(int) (d->L/2); P.jmax=(int) d->L-(int) (d->L/2); break;
    case FFTWPP_KIND_HERMITIAN:
      P.jmin=-((int) d->Lin-1); P.jmax=(int) d->Lin; break;
  }

  P.nmask=(d->N < (1ull << 31) && (d->N & (d->N-1)) == 0) ?
    (unsigned) (d->N-1) : 0u;
  {
    unsigned long long jabs=std::max<long long>(-(long long) P.jmin,P.jmax);
    P.small32=(d->N*jabs < (1ull << 32)) ? 1 : 0;
  }
  int rc;
  // zeta_N table: single level when small, else two-level with B=2^zshift
  if(d->N <= (1u << 16)) {
    P.zshift=-1;
    rc=upload_roots(d->N,d->N,&P.z1,pl->owned);
    if(rc) {delete pl; return rc;}
    P.z2=NULL;
  } else {
    int sh=0;
    while((1ull << (2*sh)) < d->N) ++sh;
    P.zshift=sh;
    size_t Bz=(size_t) 1 << sh;
    size_t nhi=(d->N+Bz-1)/Bz;
    rc=upload_roots(nhi,d->N,&P.z1,pl->owned,Bz);
    if(rc) {delete pl; return rc;}
    rc=upload_roots(Bz,d->N,&P.z2,pl->owned);
    if(rc) {delete pl; return rc;}
  }

  // distinct FFT lengths among the sub-blocks (at most 2: m and m/2)
  int ntab=0;
  pl->mmax=0;
  std::vector<SubBlockDev> hs(d->nsub);
  for(size_t i=0; i < d->nsub; ++i) {
    const fftwpp_gpu_subblock& s=d->sub[i];
    if(s.mlen == 0 || s.nout == 0 || s.nout > s.mlen || s.k0 >= d->N) {
      set_error("plan_create: invalid sub-block %zu",i);
      delete pl;
      return FFTWPP_GPU_EINVAL;
    }
    int it=-1;
    for(int k=0; k < ntab; ++k)
      if(P.tab[k].n == (int) s.mlen) it=k;
    if(it < 0) {
      if(ntab == 2) {
        set_error("plan_create: more than two distinct FFT lengths");
        delete pl;
        return FFTWPP_GPU_EUNSUPPORTED;
      }
      it=ntab++;
      if(!factorize((int) s.mlen,P.tab[it])) {
        set_error("plan_create: FFT length %u has a prime factor > %d",
                  s.mlen,MAXPRIME);
        delete pl;
        return FFTWPP_GPU_EUNSUPPORTED;
      }
      rc=upload_roots(s.mlen,s.mlen,&P.tab[it].omega,pl->owned);
      if(rc) {delete pl; return rc;}
      rc=upload_tw8(s.mlen,&P.tab[it].tw8,pl->owned);
      if(rc) {delete pl; return rc;}
    }
    hs[i].mlen=s.mlen;
    hs[i].nout=s.nout;
    hs[i].flags=s.flags;
    hs[i].tab=it;
    hs[i].k0=(long long) s.k0;
    hs[i].off_call=(long long) s.off_call;
    hs[i].off_all=(long long) s.off_all;
    if(s.mlen > pl->mmax) pl->mmax=s.mlen;
  }
  pl->hsub=hs;
  void *dsb=NULL;
  ce=cudaMalloc(&dsb,hs.size()*sizeof(SubBlockDev));
  if(ce != cudaSuccess) {delete pl; return cuda_fail(ce,"cudaMalloc(subblocks)");}
  pl->owned.push_back(dsb);
  ce=cudaMemcpy(dsb,hs.data(),hs.size()*sizeof(SubBlockDev),
                cudaMemcpyHostToDevice);
  if(ce != cudaSuccess) {delete pl; return cuda_fail(ce,"cudaMemcpy(subblocks)");}
  pl->dsub=(const SubBlockDev *) dsb;
  cudaGetDevice(&pl->device);
  fast_plan_init(pl);
  *out=pl;
  return FFTWPP_GPU_OK;
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------

static const size_t SMEM_BUDGET=200*1024;
static const size_t SMEM_MAX=227*1024;

template<class K>
static int enable_smem(K kernel)
{
  static std::mutex mu;
  static std::vector<const void *> done[16];
  int dev=0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  std::vector<const void *>& v=done[dev & 15];
  for(size_t i=0; i < v.size(); ++i)
    if(v[i] == (const void *) kernel) return 0;
  cudaError_t e=cudaFuncSetAttribute(kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int) SMEM_MAX);
  if(e != cudaSuccess) return cuda_fail(e,"cudaFuncSetAttribute");
  v.push_back((const void *) kernel);
  return 0;
}

static size_t inword(int kind)
{
  return kind == FFTWPP_KIND_REAL ? sizeof(double) : sizeof(double2);
}

// Choose the tile: strided passes take TC columns of one row; contiguous
// batches take TR rows.
static int choose_tile(const Plan *pl, size_t narr_in, size_t narr_w,
                       uint64_t nrows, int& TR, int& TC, size_t& inbytes,
                       size_t& wbytes)
{
  size_t per_lane=narr_in*pl->dev.Lin*inword(pl->dev.kind)+
    narr_w*pl->mmax*sizeof(double2);
  if(per_lane > SMEM_MAX) {
    set_error("transform of length m=%u, L=%d does not fit in shared memory "
              "(%zu bytes per lane); choose a smaller m (inner scheme)",
              pl->mmax,pl->dev.Lin,per_lane);
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  size_t lanes=SMEM_BUDGET/per_lane;
  if(lanes < 1) lanes=1;
  if(pl->dev.C > 1) {
    TR=1;
    TC=(int) std::min<size_t>(std::min<size_t>(lanes,8),pl->dev.C);
  } else {
    TC=1;
    size_t want=std::max<size_t>(1,(4*NTHREADS)/pl->mmax);
    TR=(int) std::min<size_t>(std::min<size_t>(lanes,std::max<size_t>(want,1)),
                              std::max<uint64_t>(nrows,1));
    if(TR > 16) TR=16;
  }
  size_t T=(size_t) TR*TC;
  inbytes=(pl->dev.Lin*T*inword(pl->dev.kind)+15) & ~(size_t) 15;
  wbytes=pl->mmax*T*sizeof(double2);
  return 0;
}

// ---- optional profiling: one event pair per launch, summed per key ----
struct ProfRec {int key; cudaEvent_t e0, e1;};
static bool g_prof=false;
static std::vector<ProfRec> g_prof_recs;
static int g_prof_open=-1;
static std::mutex g_prof_mu;

void prof_begin(int key, cudaStream_t st)
{
  if(!g_prof) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  ProfRec r;
  r.key=key;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0,st);
  g_prof_recs.push_back(r);
  g_prof_open=(int) g_prof_recs.size()-1;
}

// SMs of the current device (grids are sized in multiples of it)
int sm_count()
{
  static int n[16]={0};
  int dev=0;
  cudaGetDevice(&dev);
  int& v=n[dev & 15];
  if(v == 0) {
    if(cudaDeviceGetAttribute(&v,cudaDevAttrMultiProcessorCount,dev) !=
       cudaSuccess || v <= 0)
      v=148;
  }
  return v;
}

int check_launch(const char *what, cudaStream_t st)
{
  cudaError_t e=cudaGetLastError();
  if(g_prof && g_prof_open >= 0) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    cudaEventRecord(g_prof_recs[g_prof_open].e1,st);
    g_prof_open=-1;
  }
  if(e != cudaSuccess) return cuda_fail(e,what);
  g_launches.fetch_add(1);
  return 0;
}

int prof_enable(int on)
{
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for(size_t i=0; i < g_prof_recs.size(); ++i) {
    cudaEventDestroy(g_prof_recs[i].e0);
    cudaEventDestroy(g_prof_recs[i].e1);
  }
  g_prof_recs.clear();
  g_prof_open=-1;
  g_prof=on != 0;
  return 0;
}

int prof_read(double *ms, uint64_t *count)
{
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for(int k=0; k < PROF_KEYS; ++k) {ms[k]=0.0; count[k]=0;}
  cudaError_t e=cudaDeviceSynchronize();
  if(e != cudaSuccess) return cuda_fail(e,"cudaDeviceSynchronize");
  for(size_t i=0; i < g_prof_recs.size(); ++i) {
    float t=0.0f;
    if(cudaEventElapsedTime(&t,g_prof_recs[i].e0,g_prof_recs[i].e1) == cudaSuccess) {
      int k=g_prof_recs[i].key & (PROF_KEYS-1);
      ms[k] += t;
      ++count[k];
    }
  }
  return 0;
}

#define DISPATCH_KIND(kind, ...)                                    \
  switch(kind) {                                                    \
    case FFTWPP_KIND_COMPLEX: {const int K=FFTWPP_KIND_COMPLEX; __VA_ARGS__; break;}     \
    case FFTWPP_KIND_CENTERED: {const int K=FFTWPP_KIND_CENTERED; __VA_ARGS__; break;}   \
    case FFTWPP_KIND_HERMITIAN: {const int K=FFTWPP_KIND_HERMITIAN; __VA_ARGS__; break;} \
    default: {const int K=FFTWPP_KIND_REAL; __VA_ARGS__; break;}                         \
  }

int generic_forward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                    const void *f, void *F, uint64_t nrows, uint64_t frs,
                    uint64_t Frs, cudaStream_t st)
{
  int TR,TC;
  size_t inbytes,wbytes;
  int rc=choose_tile(pl,1,1,nrows,TR,TC,inbytes,wbytes);
  if(rc) return rc;
  int ntc=(pl->dev.C+TC-1)/TC;
  uint64_t ntr=(nrows+TR-1)/TR;
  uint64_t grid=ntr*ntc;
  if(grid == 0) return 0;
  if(grid > 0x7fffffffull) {
    set_error("forward: grid too large");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  size_t smem=inbytes+wbytes;
  DISPATCH_KIND(pl->dev.kind,
    rc=enable_smem(forward_kernel<K>);
    if(rc) return rc;
    prof_begin(4*pl->tag+0,st);
    forward_kernel<K><<<(unsigned) grid,NTHREADS,smem,st>>>
      (pl->dev,pl->dsub+sb0,(int) nsb,layout,f,F,(long long) nrows,
       (long long) frs,(long long) Frs,TR,TC,ntc,inbytes));
  return check_launch("forward_kernel",st);
}

int generic_backward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                     const void *F, void *f, int accumulate, double scale,
                     uint64_t nrows, uint64_t Frs, uint64_t frs,
                     cudaStream_t st)
{
  int TR,TC;
  size_t accbytes,wbytes;
  int rc=choose_tile(pl,1,1,nrows,TR,TC,accbytes,wbytes);
  if(rc) return rc;
  int ntc=(pl->dev.C+TC-1)/TC;
  uint64_t ntr=(nrows+TR-1)/TR;
  uint64_t grid=ntr*ntc;
  if(grid == 0) return 0;
  if(grid > 0x7fffffffull) {
    set_error("backward: grid too large");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  size_t smem=accbytes+wbytes;
  DISPATCH_KIND(pl->dev.kind,
    rc=enable_smem(backward_kernel<K>);
    if(rc) return rc;
    prof_begin(4*pl->tag+1,st);
    backward_kernel<K><<<(unsigned) grid,NTHREADS,smem,st>>>
      (pl->dev,pl->dsub+sb0,(int) nsb,layout,F,f,accumulate,scale,
       (long long) nrows,(long long) Frs,(long long) frs,TR,TC,ntc,accbytes));
  return check_launch("backward_kernel",st);
}

int generic_convolve(Plan *pl, void *const *f, uint32_t A, uint32_t B,
                     int mult, double scale, uint64_t nrows, uint64_t rs,
                     cudaStream_t st)
{
  if(pl->dev.C != 1) {
    set_error("convolve: plan must have C == 1");
    return FFTWPP_GPU_EINVAL;
  }
  int TR,TC;
  size_t inbytes,wbytes;
  uint32_t nw=std::max(A,B);
  int rc=choose_tile(pl,A+B,nw,nrows,TR,TC,inbytes,wbytes);
  if(rc) return rc;
  uint64_t grid=(nrows+TR-1)/TR;
  if(grid == 0) return 0;
  if(grid > 0x7fffffffull) {
    set_error("convolve: grid too large");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  ConvPtrs ptrs;
  for(uint32_t a=0; a < nw; ++a) ptrs.p[a]=f[a];
  size_t wstride=(size_t) pl->mmax*TR;
  size_t smem=(size_t) (A+B)*inbytes+(size_t) nw*wbytes;
  DISPATCH_KIND(pl->dev.kind,
    rc=enable_smem(convolve_kernel<K>);
    if(rc) return rc;
    prof_begin(4*pl->tag+2,st);
    convolve_kernel<K><<<(unsigned) grid,NTHREADS,smem,st>>>
      (pl->dev,pl->dsub,(int) pl->hsub.size(),ptrs,(int) A,(int) B,mult,
       scale,(long long) nrows,(long long) rs,TR,inbytes,wstride));
  return check_launch("convolve_kernel",st);
}

int launch_scale(double *x, double scale, uint64_t n0, uint64_t n1,
                 uint64_t n2, uint64_t s0, uint64_t s1, cudaStream_t st)
{
  uint64_t total=n0*n1*n2;
  if(total == 0) return 0;
  unsigned grid=(unsigned) std::min<uint64_t>((total+255)/256,(uint64_t) sm_count()*16);
  prof_begin(3,st);
  scale_kernel<<<grid,256,0,st>>>(x,scale,n0,n1,n2,s0,s1);
  return check_launch("scale_kernel",st);
}

int launch_copy3(void *dst, const void *src, uint64_t n0, uint64_t n1,
                 uint64_t n2, uint64_t d0, uint64_t d1, uint64_t s0,
                 uint64_t s1, cudaStream_t st)
{
  uint64_t total=n0*n1*n2;
  if(total == 0) return 0;
  unsigned grid=(unsigned) std::min<uint64_t>((total+255)/256,(uint64_t) sm_count()*16);
  prof_begin(3,st);
  copy3_kernel<<<grid,256,0,st>>>((double2 *) dst,(const double2 *) src,
                                   n0,n1,n2,d0,d1,s0,s1);
  return check_launch("copy3_kernel",st);
}

} // namespace fftwpp_gpu
