// fast_kernels.cu -- specialised sm_100a kernels for power-of-two residue
// passes (the shapes of every BASELINE config).  Same mathematics and same
// C ABI as the generic kernels in gpu_core.cu; what changes is the execution:
//
//  * every thread owns 8 complex points in registers; an FFT of length
//    N = 2^k is a sequence of register radix-8 butterflies (plus one radix-2/4
//    pass when 3 does not divide k) with shared-memory exchanges in between --
//    decimation in frequency forward (digit-reversed result), exact adjoint
//    backward, so the fused convolution multiplies in scrambled order and never
//    reorders;
//  * implicit zero padding, residue twiddles, the multiplier and the
//    conjugate-twiddle accumulation over residues all happen in registers: the
//    padded data never exists in HBM and the accumulators of the fused
//    convolution never leave the register file;
//  * strided ("Many") passes take tiles of T adjacent columns so that every
//    global access is a run of T*16 contiguous bytes, and keep the tile
//    lane-fastest in shared memory (bank-conflict free for every stage);
//    contiguous rows use a padded per-row exchange buffer (index p + p/8)
//    and per-row named barriers so rows of one CTA run independently.
//
// Reference routines covered: fftPad::forward1/forward2[Many] + backward
// (convolve.cc:849-1225,1482-1763), fftPadReal::forward1Many/backward1Many
// (convolve.cc:5852-5964,6702-6788), the residue loop of
// Convolution::convolveRaw with multBinary/multcorrelation fused
// (convolve.cc:7513-7575,33-110).

#include "regfft.cuh"

#include <mutex>

namespace fftwpp_gpu {

namespace {

template<int KIND>
struct Word {typedef double2 type;};
template<>
struct Word<FFTWPP_KIND_REAL> {typedef double type;};


// ---------------------------------------------------------------------------
// fused 1-D convolution over contiguous rows (COMPLEX kind, A=2, B=1)
// ---------------------------------------------------------------------------

// Persistent CTAs: each loads the twiddle tables (radix-8 twiddles and the
// residue twiddles zeta^{k0 j} of every sub-block with k0 != 0) into shared
// memory once and then loops over groups of ROWS rows.  Per row: two padded
// exchange buffers; the FFTs of the two inputs and the inverse FFT run
// through buffer 0, the running accumulators rest in buffer 1 meanwhile, so
// the kernel fits 128 registers (2 CTAs/SM) without local-memory spills.
template<int LG, int NTERM>
__global__ void __launch_bounds__((1 << LG)/8 > 256 ? (1 << LG)/8 : 256, (1 << LG)/8 > 256 ? 1 : 2)
fast_conv_rows(PlanDev P, const SubBlockDev *__restrict__ sbs, int nsb,
               double2 *f0, const double2 *f1, int mult, double scale,
               long long nrows, long long rs, int tabid, int zlen,
               long long ngroups)
{
  typedef RegFFT<LG> FFT;
  const int M=FFT::N;
  const int TPT=FFT::TPT;
  const int NT=TPT > 256 ? TPT : 256;
  const int ROWS=NT/TPT;
  const int TWN=FFT::twCount();
  const int BUF=NTERM*M+NTERM*M/8; // also holds NTERM*M parked accumulators
  extern __shared__ __align__(16) double2 sm[];
  double2 *tws=sm;
  double2 *zs=sm+TWN;               // zlen entries per sub-block with k0 != 0
  int nz=0;
  for(int isb=0; isb < nsb; ++isb) nz += sbs[isb].k0 != 0;
  double2 *bufs=zs+(zlen ? (size_t) nz*zlen : 0);
  const int rowInCta=threadIdx.x/TPT;
  const int tau=threadIdx.x % TPT;
  const int L=P.jmax;

  // tables -> shared memory
  {
    const double2 *tw=P.tab[tabid].tw8;
    for(int i=threadIdx.x; i < TWN; i += NT) tws[i]=__ldg(tw+i);
    if(zlen) {
      int slot=0;
      for(int isb=0; isb < nsb; ++isb) {
        const long long k0=sbs[isb].k0;
        if(k0 == 0) continue;
        for(int j=threadIdx.x; j < zlen; j += NT)
          zs[(size_t) slot*zlen+j]=zeta(P,modN(P,k0,j));
        ++slot;
      }
    }
  }
  __syncthreads();

  RowLayout lay;
  lay.base=rowInCta*2*BUF;
  lay.barid=TPT > 32 ? 1+rowInCta : 0;
  lay.nthreads=TPT;
  double2 *park=bufs+lay.base+BUF+tau;

  for(long long grp=blockIdx.x; grp < ngroups; grp += gridDim.x) {
    long long row=grp*ROWS+rowInCta;
    const bool live=row < nrows;
    if(!live) row=nrows-1;
    double2 *g0=f0+row*rs;
    const double2 *g1=f1+row*rs;
    if(P.C == 1) { // (always) pull the next group's rows into L2 ahead of use
      long long nrow=(grp+gridDim.x)*ROWS+rowInCta;
      if(nrow < nrows) {
        const char *n0=(const char *) (f0+nrow*rs);
        const char *n1=(const char *) (f1+nrow*rs);
        for(int off=tau*128; off < L*16; off += TPT*128) {
          asm volatile("prefetch.global.L2 [%0];" :: "l"(n0+off));
          asm volatile("prefetch.global.L2 [%0];" :: "l"(n1+off));
        }
      }
    }

    double2 acc[NTERM][8];
#pragma unroll
    for(int k=0; k < NTERM; ++k)
#pragma unroll
      for(int t=0; t < 8; ++t)
        acc[k][t]=make_double2(0.0,0.0);

    int slot=0;
    for(int isb=0; isb < nsb; ++isb) {
      const long long k0=sbs[isb].k0;
      const double2 *zrow=zs+(size_t) slot*zlen;
      if(k0 != 0) ++slot;
      if(isb > 0) {
#pragma unroll
        for(int k=0; k < NTERM; ++k)
#pragma unroll
          for(int t=0; t < 8; ++t)
            park[(k*8+t)*TPT]=acc[k][t];
      }
      double2 x[1][8], y[1][8];
#pragma unroll
      for(int t=0; t < 8; ++t) {
        x[0][t]=make_double2(0.0,0.0);
        y[0][t]=make_double2(0.0,0.0);
#pragma unroll
        for(int k=0; k < NTERM; ++k) {
          int j=tau+TPT*t+k*M;
          if(j < L) {
            double2 a=g0[j];
            double2 b=g1[j];
            if(k0 != 0) {
              double2 z=zlen ? zrow[j] : zeta(P,modN(P,k0,j));
              a=fmul(a,z);
              b=fmul(b,z);
            }
            x[0][t]=x[0][t]+a;
            y[0][t]=y[0][t]+b;
          }
        }
      }
      FFT::template forward<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
      FFT::template forward<1,RowLayout,true,false,2>(y,tau,tws,bufs,0,lay,true);
      if(mult == FFTWPP_MULT_BINARY) {
#pragma unroll
        for(int t=0; t < 8; ++t) x[0][t]=fmul(x[0][t],y[0][t]);
      } else {
#pragma unroll
        for(int t=0; t < 8; ++t) x[0][t]=fmulc(x[0][t],y[0][t]);
      }
      FFT::template adjoint<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
      if(isb > 0) {
#pragma unroll
        for(int k=0; k < NTERM; ++k)
#pragma unroll
          for(int t=0; t < 8; ++t)
            acc[k][t]=park[(k*8+t)*TPT];
      }
#pragma unroll
      for(int t=0; t < 8; ++t) {
#pragma unroll
        for(int k=0; k < NTERM; ++k) {
          int j=tau+TPT*t+k*M;
          if(j < L) {
            double2 v=x[0][t];
            if(k0 != 0)
              v=fmulc(v,zlen ? zrow[j] : zeta(P,modN(P,k0,j)));
            acc[k][t]=acc[k][t]+v;
          }
        }
      }
    }
    if(live) {
#pragma unroll
      for(int k=0; k < NTERM; ++k)
#pragma unroll
        for(int t=0; t < 8; ++t) {
          int j=tau+TPT*t+k*M;
          if(j < L)
            g0[j]=make_double2(acc[k][t].x*scale,acc[k][t].y*scale);
        }
    }
    // the next group's first exchange starts with a barrier, which also
    // orders this group's last reads of the exchange buffers
  }
}

// Software-pipelined variant for L <= m (one input term per W[s], the shape
// of every p=1 configuration): the kernel is bound by the latency of its
// global loads at 16 warps/SM, so each work item (row, sub-block) issues the
// loads of the NEXT item's second input right after the multiplier, where
// those registers fall free, and the loads of its own first input just before
// the FFT of the second -- both land while an FFT runs.  The running sum over
// sub-blocks rests in the idle half of the row's shared-memory buffer between
// items, so no accumulator registers are live across the FFTs.
template<int LG>
__global__ void __launch_bounds__((1 << LG)/8 > 256 ? (1 << LG)/8 : 256, (1 << LG)/8 > 256 ? 1 : 2)
fast_conv_rows_pipe(PlanDev P, const SubBlockDev *__restrict__ sbs, int nsb,
                    double2 *f0, const double2 *f1, int mult, double scale,
                    long long nrows, long long rs, int tabid, int zlen,
                    long long ngroups)
{
  typedef RegFFT<LG> FFT;
  const int M=FFT::N;
  const int TPT=FFT::TPT;
  const int NT=TPT > 256 ? TPT : 256;
  const int ROWS=NT/TPT;
  const int TWN=FFT::twCount();
  const int BUF=M+M/8;
  extern __shared__ __align__(16) double2 sm[];
  double2 *tws=sm;
  double2 *zs=sm+TWN;
  int nz=0;
  for(int isb=0; isb < nsb; ++isb) nz += sbs[isb].k0 != 0;
  double2 *bufs=zs+(zlen ? (size_t) nz*zlen : 0);
  const int rowInCta=threadIdx.x/TPT;
  const int tau=threadIdx.x % TPT;
  const int L=P.jmax;

  {
    const double2 *tw=P.tab[tabid].tw8;
    for(int i=threadIdx.x; i < TWN; i += NT) tws[i]=__ldg(tw+i);
    if(zlen) {
      int slot=0;
      for(int isb=0; isb < nsb; ++isb) {
        const long long k0=sbs[isb].k0;
        if(k0 == 0) continue;
        for(int j=threadIdx.x; j < zlen; j += NT)
          zs[(size_t) slot*zlen+j]=zeta(P,modN(P,k0,j));
        ++slot;
      }
    }
  }
  __syncthreads();

  RowLayout lay;
  lay.base=rowInCta*2*BUF;
  lay.barid=TPT > 32 ? 1+rowInCta : 0;
  lay.nthreads=TPT;
  double2 *park=bufs+lay.base+BUF+tau;

  long long grp=blockIdx.x;
  if(grp >= ngroups) return;
  double2 x[1][8], y[1][8];
  {
    long long row=grp*ROWS+rowInCta;
    if(row >= nrows) row=nrows-1;
    const double2 *g1=f1+row*rs;
#pragma unroll
    for(int t=0; t < 8; ++t) {
      const int j=tau+TPT*t;
      y[0][t]=j < L ? g1[j] : make_double2(0.0,0.0);
    }
  }

  for(; grp < ngroups; grp += gridDim.x) {
    long long row=grp*ROWS+rowInCta;
    const bool live=row < nrows;
    if(!live) row=nrows-1;
    double2 *g0=f0+row*rs;
    const double2 *g1=f1+row*rs;
    const long long ngrp=grp+gridDim.x;
    const bool more=ngrp < ngroups;
    long long nrow=ngrp*ROWS+rowInCta;
    if(nrow >= nrows) nrow=nrows-1;
    const double2 *n1=f1+nrow*rs;
    if(more) { // pull the next group's rows into L2 ahead of their register loads
      const char *p0=(const char *) (f0+nrow*rs);
      const char *p1=(const char *) n1;
      for(int off=tau*128; off < L*16; off += TPT*128) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p0+off));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p1+off));
      }
    }

    int slot=0;
    for(int isb=0; isb < nsb; ++isb) {
      const long long k0=sbs[isb].k0;
      const double2 *zrow=zs+(size_t) slot*zlen;
      if(k0 != 0) ++slot;
      // this item's first input: in flight during the FFT of the second
#pragma unroll
      for(int t=0; t < 8; ++t) {
        const int j=tau+TPT*t;
        x[0][t]=j < L ? g0[j] : make_double2(0.0,0.0);
      }
      if(k0 != 0) {
#pragma unroll
        for(int t=0; t < 8; ++t) {
          const int j=tau+TPT*t;
          if(j < L)
            y[0][t]=fmul(y[0][t],zlen ? zrow[j] : zeta(P,modN(P,k0,j)));
        }
      }
      FFT::template forward<1,RowLayout,true,false,2>(y,tau,tws,bufs,0,lay,true);
      if(k0 != 0) {
#pragma unroll
        for(int t=0; t < 8; ++t) {
          const int j=tau+TPT*t;
          if(j < L)
            x[0][t]=fmul(x[0][t],zlen ? zrow[j] : zeta(P,modN(P,k0,j)));
        }
      }
      FFT::template forward<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
      if(mult == FFTWPP_MULT_BINARY) {
#pragma unroll
        for(int t=0; t < 8; ++t) x[0][t]=fmul(x[0][t],y[0][t]);
      } else {
#pragma unroll
        for(int t=0; t < 8; ++t) x[0][t]=fmulc(x[0][t],y[0][t]);
      }
      // the next item's second input: in flight during the inverse FFT
      const bool lastsb=isb+1 == nsb;
      if(!lastsb || more) {
        const double2 *h1=lastsb ? n1 : g1;
#pragma unroll
        for(int t=0; t < 8; ++t) {
          const int j=tau+TPT*t;
          y[0][t]=j < L ? h1[j] : make_double2(0.0,0.0);
        }
      }
      FFT::template adjoint<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
#pragma unroll
      for(int t=0; t < 8; ++t) {
        const int j=tau+TPT*t;
        if(j < L) {
          double2 v=x[0][t];
          if(k0 != 0)
            v=fmulc(v,zlen ? zrow[j] : zeta(P,modN(P,k0,j)));
          if(isb > 0) v=v+park[t*TPT];
          if(!lastsb) park[t*TPT]=v;
          else if(live) g0[j]=make_double2(v.x*scale,v.y*scale);
        }
      }
    }
  }
}

// The same software pipeline specialised for the shape of every BASELINE z
// pass: fftPad with p=1, q=2 (sub-blocks k0=0 and k0=1) and full rows (L == m).
// No per-point bounds or residue predicates, and the residue twiddles are not
// read from a table per point: zeta^{tau+TPT t} = zeta^tau (zeta^TPT)^t by
// successive products from one shared-memory word per thread and a kernel
// constant (the shared-memory pipe, not FP64 issue, is the busier one: this
// removes 21 of the 24 128-bit residue-twiddle reads per row).
template<int LG, int MULT>
__global__ void __launch_bounds__((1 << LG)/8 > 256 ? (1 << LG)/8 : 256, (1 << LG)/8 > 256 ? 1 : 2)
fast_conv_rows_q2(PlanDev P, const SubBlockDev *__restrict__ sbs,
                  double2 *f0, const double2 *f1, double scale,
                  const double2 zstep, long long nrows, long long rs,
                  int tabid, long long ngroups)
{
  typedef WFFT<LG> FFT;
  typedef RegFFT<LG> RF;
  const int M=1 << LG;
  const int TPT=M/8;
  const int NT=TPT > 256 ? TPT : 256;
  const int ROWS=NT/TPT;
  const int BUF=M+M/8;
  const int TWN=RF::twCount();
  extern __shared__ __align__(16) double2 sm[];
  double2 *tws=sm;
  double2 *bufs=sm+TWN;
  const int rowInCta=threadIdx.x/TPT;
  const int tau=threadIdx.x % TPT;
  for(int i=threadIdx.x; i < TWN; i += NT) tws[i]=__ldg(P.tab[tabid].tw8+i);
  __syncthreads();
  double2 *zs=bufs;
  bufs += TPT;
  for(int i=threadIdx.x; i < TPT; i += NT) zs[i]=zeta(P,modN(P,sbs[1].k0,i));
  __syncthreads();

  RowLayout lay;
  lay.base=rowInCta*2*BUF;
  lay.barid=TPT > 32 ? 1+rowInCta : 0;
  lay.nthreads=TPT;
  double2 *park=bufs+lay.base+BUF+tau;

  long long grp=blockIdx.x;
  if(grp >= ngroups) return;
  double2 x[1][8], y[1][8];
  {
    long long row=grp*ROWS+rowInCta;
    if(row >= nrows) row=nrows-1;
    const double2 *g1=f1+row*rs+tau;
#pragma unroll
    for(int t=0; t < 8; ++t) y[0][t]=g1[TPT*t];
  }
  for(; grp < ngroups; grp += gridDim.x) {
    long long row=grp*ROWS+rowInCta;
    const bool live=row < nrows;
    if(!live) row=nrows-1;
    double2 *g0=f0+row*rs+tau;
    const double2 *g1=f1+row*rs+tau;
    const long long ngrp=grp+gridDim.x;
    const bool more=ngrp < ngroups;
    long long nrow=ngrp*ROWS+rowInCta;
    if(nrow >= nrows) nrow=nrows-1;
    const double2 *n1=f1+nrow*rs+tau;
    if(more) {
      const char *p0=(const char *) (f0+nrow*rs);
      const char *p1=(const char *) (f1+nrow*rs);
      for(int off=tau*128; off < M*16; off += TPT*128) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p0+off));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p1+off));
      }
    }
#pragma unroll 1
    for(int isb=0; isb < 2; ++isb) {
      const bool hz=isb == 1;
#pragma unroll
      for(int t=0; t < 8; ++t) x[0][t]=g0[TPT*t];
      if(hz) {
        double2 z=zs[tau];
#pragma unroll
        for(int t=0; t < 8; ++t) {
          y[0][t]=fmul(y[0][t],z);
          if(t < 7) z=fmul(z,zstep);
        }
      }
      RF::template forward<1,RowLayout,true,false,2>(y,tau,tws,bufs,0,lay,true);
      if(hz) {
        double2 z=zs[tau];
#pragma unroll
        for(int t=0; t < 8; ++t) {
          x[0][t]=fmul(x[0][t],z);
          if(t < 7) z=fmul(z,zstep);
        }
      }
      RF::template forward<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
#pragma unroll
      for(int t=0; t < 8; ++t)
        x[0][t]=MULT == FFTWPP_MULT_BINARY ? fmul(x[0][t],y[0][t]) : fmulc(x[0][t],y[0][t]);
      if(!hz || more) {
        const double2 *h1=hz ? n1 : g1;
#pragma unroll
        for(int t=0; t < 8; ++t) y[0][t]=h1[TPT*t];
      }
      RF::template adjoint<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
      if(!hz) {
#pragma unroll
        for(int t=0; t < 8; ++t) park[t*TPT]=x[0][t];
      } else {
        double2 z=zs[tau];
#pragma unroll
        for(int t=0; t < 8; ++t) {
          double2 v=fmulc(x[0][t],z)+park[t*TPT];
          if(t < 7) z=fmul(z,zstep);
          if(live) g0[TPT*t]=make_double2(v.x*scale,v.y*scale);
        }
      }
    }
  }
}


// Hermitian rows (fftPadHermitian p=2 or explicit; reference forward2/
// backward2, convolve.cc:4517-4609,4749-4843, with realMultBinary): the input
// holds the H non-negative modes of a real signal, so every residue's
//   W[s] = zeta^{k0 s} f[s] + zeta^{k0 (s-m)} conj(f[m-s])
// is Hermitian in s and its transform is REAL.  The two inputs therefore share
// ONE complex FFT (Z = W_0 + i W_1 -> X_0 + i X_1), the multiplier is
// Re(Z)*Im(Z), and one adjoint FFT of that real sequence returns the residue's
// contribution: 2 FFTs per residue instead of 3.
template<int LG>
__global__ void __launch_bounds__((1 << LG)/8 > 256 ? (1 << LG)/8 : 256, (1 << LG)/8 > 256 ? 1 : 2)
fast_conv_rows_herm(PlanDev P, const SubBlockDev *__restrict__ sbs, int nsb,
                    double2 *f0, const double2 *f1, double scale,
                    long long nrows, long long rs, int tabid, int zlen,
                    long long ngroups)
{
  typedef RegFFT<LG> FFT;
  const int M=FFT::N;
  const int TPT=FFT::TPT;
  const int NT=TPT > 256 ? TPT : 256;
  const int ROWS=NT/TPT;
  const int TWN=FFT::twCount();
  const int BUF=M+M/8;
  extern __shared__ __align__(16) double2 sm[];
  double2 *tws=sm;
  double2 *zs=sm+TWN;   // zlen=2H-1 entries (j=-(H-1)..H-1) per k0 != 0 slot
  int nz=0;
  for(int isb=0; isb < nsb; ++isb) nz += sbs[isb].k0 != 0;
  double2 *bufs=zs+(zlen ? (size_t) nz*zlen : 0);
  const int rowInCta=threadIdx.x/TPT;
  const int tau=threadIdx.x % TPT;
  const int H=P.jmax; // stored modes 0..H-1; jmin=-(H-1)
  {
    const double2 *tw=P.tab[tabid].tw8;
    for(int i=threadIdx.x; i < TWN; i += NT) tws[i]=__ldg(tw+i);
    if(zlen) {
      int slot=0;
      for(int isb=0; isb < nsb; ++isb) {
        const long long k0=sbs[isb].k0;
        if(k0 == 0) continue;
        for(int j=threadIdx.x; j < zlen; j += NT)
          zs[(size_t) slot*zlen+j]=zeta(P,modN(P,k0,j+P.jmin));
        ++slot;
      }
    }
  }
  __syncthreads();

  RowLayout lay;
  lay.base=rowInCta*2*BUF;
  lay.barid=TPT > 32 ? 1+rowInCta : 0;
  lay.nthreads=TPT;
  double2 *park=bufs+lay.base+BUF+tau;

  for(long long grp=blockIdx.x; grp < ngroups; grp += gridDim.x) {
    long long row=grp*ROWS+rowInCta;
    const bool live=row < nrows;
    if(!live) row=nrows-1;
    double2 *g0=f0+row*rs;
    const double2 *g1=f1+row*rs;

    double2 acc[8];
#pragma unroll
    for(int t=0; t < 8; ++t) acc[t]=make_double2(0.0,0.0);

    int slot=0;
    for(int isb=0; isb < nsb; ++isb) {
      const long long k0=sbs[isb].k0;
      const double2 *zrow=zs+(size_t) slot*zlen-P.jmin; // indexed by j
      if(k0 != 0) ++slot;
      if(isb > 0) {
#pragma unroll
        for(int t=0; t < 8; ++t) park[t*TPT]=acc[t];
      }
      double2 x[1][8];
#pragma unroll
      for(int t=0; t < 8; ++t) {
        const int s=tau+TPT*t;
        double2 wa=make_double2(0.0,0.0), wb=make_double2(0.0,0.0);
        if(s < H) {
          double2 a=g0[s], b=g1[s];
          if(s == 0) {a.y=0.0; b.y=0.0;} // c2r ignores Im f[0]
          if(k0 != 0) {
            double2 z=zlen ? zrow[s] : zeta(P,modN(P,k0,s));
            a=fmul(a,z);
            b=fmul(b,z);
          }
          wa=a;
          wb=b;
        }
        const int jn=M-s; // mode -(M-s)
        if(s >= 1 && jn < H) {
          double2 a=g0[jn], b=g1[jn];
          a.y=-a.y;
          b.y=-b.y;
          if(k0 != 0) {
            double2 z=zlen ? zrow[-jn] : zeta(P,modN(P,k0,-jn));
            a=fmul(a,z);
            b=fmul(b,z);
          }
          wa=wa+a;
          wb=wb+b;
        }
        // Z = W_0 + i W_1
        x[0][t]=make_double2(wa.x-wb.y,wa.y+wb.x);
      }
      FFT::template forward<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
#pragma unroll
      for(int t=0; t < 8; ++t)
        x[0][t]=make_double2(x[0][t].x*x[0][t].y,0.0); // realMultBinary
      FFT::template adjoint<1,RowLayout,true,false,2>(x,tau,tws,bufs,0,lay,true);
      if(isb > 0) {
#pragma unroll
        for(int t=0; t < 8; ++t) acc[t]=park[t*TPT];
      }
#pragma unroll
      for(int t=0; t < 8; ++t) {
        const int s=tau+TPT*t;
        if(s < H) {
          double2 v=x[0][t];
          if(k0 != 0)
            v=fmulc(v,zlen ? zrow[s] : zeta(P,modN(P,k0,s)));
          acc[t]=acc[t]+v;
        }
      }
    }
    if(live) {
#pragma unroll
      for(int t=0; t < 8; ++t) {
        const int s=tau+TPT*t;
        if(s < H)
          g0[s]=make_double2(acc[t].x*scale,acc[t].y*scale);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// strided ("Many") forward / backward passes over tiles of T columns
// ---------------------------------------------------------------------------
//
// Persistent CTAs loop over (row, column-tile) work items.  Shared memory:
//   [tw8 of length M][tw8 of length M/2 (mixed plans)][zeta rows][tile][exchange]
// DIRECT variants (uniform COMPLEX plans with L <= M): the thread's 8 input
// points are loaded straight from global memory into registers and reused for
// every residue; the backward accumulators live in registers too.  Otherwise
// the input tile / accumulator tile is staged in shared memory (needed when
// sub-blocks have different lengths: fftPadReal's packed residue class).

struct ManyTables {
  const double2 *tw[2];  // smem radix-8 twiddles for lengths M and M/2
  const double2 *zs;     // smem residue twiddles (zlen per sub-block slot)
  int zlen;
};

// logical sample j of lane `lane` from the staged tile (lane fastest)
// rs: words between logical rows (T for the staged tile, the plan stride S
// when the "tile" is read straight from global memory)
template<int KIND>
__device__ __forceinline__ double2 tileInput(const void *in, int j, int jmin,
                                             long long rs, int lane)
{
  if(KIND == FFTWPP_KIND_REAL)
    return make_double2(((const double *) in)[j*rs+lane],0.0);
  const double2 *c=(const double2 *) in;
  if(KIND == FFTWPP_KIND_HERMITIAN) {
    if(j >= 0) return c[j*rs+lane];
    double2 v=c[(-j)*rs+lane];
    return make_double2(v.x,-v.y);
  }
  return c[(j-jmin)*rs+lane];
}

__device__ __forceinline__ double2 zetaAt(const PlanDev& P,
                                          const ManyTables& tb, int slot,
                                          long long k0, int j)
{
  // table rows are indexed by the storage index j-jmin
  if(tb.zlen) return tb.zs[(size_t) slot*tb.zlen+(j-P.jmin)];
  return zeta(P,modN(P,k0,j));
}

struct NoPrefetch {
  __device__ __forceinline__ void operator()() const {}
};

// `pre` runs after the thread's inputs have been consumed and before the FFT:
// the DIRECT kernels use it to put the next tile's loads in flight.
template<int KIND, int LG, bool DIRECT, class Pre>
__device__ __forceinline__ void forwardSub(const PlanDev& P,
                                           const SubBlockDev& sb, int slot,
                                           const ManyTables& tb, int which,
                                           const void *in,
                                           const double2 (&xin)[8],
                                           double2 *buf, void *F,
                                           long long Fbase, int T, int col0,
                                           bool colsok, long long plane,
                                           long long rs, int& pp,
                                           int ppStride, const Pre& pre)
{
  typedef RegFFT<LG> FFT;
  const int mlen=FFT::N;
  const int TPT=FFT::TPT;
  const int lane=threadIdx.x % T;
  const int tau=threadIdx.x/T;
  const bool active=tau < TPT;
  LaneLayout lay;
  lay.T=T;
  lay.lane=lane;
  const long long k0=sb.k0;
  double2 x[1][8];
#pragma unroll
  for(int t=0; t < 8; ++t) x[0][t]=make_double2(0.0,0.0);
  if(active) {
    if(DIRECT) {
#pragma unroll
      for(int t=0; t < 8; ++t) {
        int j=tau+TPT*t;
        double2 v=xin[t];
        if(k0 != 0 && j < P.jmax) v=fmul(v,zetaAt(P,tb,slot,k0,j));
        x[0][t]=v;
      }
    } else {
#pragma unroll
      for(int t=0; t < 8; ++t) {
        int s=tau+TPT*t;
        int d=(s-P.jmin) & (mlen-1);
        double2 acc=make_double2(0.0,0.0);
        for(int j=P.jmin+d; j < P.jmax; j += mlen) {
          double2 v=(rs == T || colsok) ? tileInput<KIND>(in,j,P.jmin,rs,lane)
            : make_double2(0.0,0.0);
          if(k0 != 0) v=fmul(v,zetaAt(P,tb,slot,k0,j));
          acc=acc+v;
        }
        x[0][t]=acc;
      }
    }
  }
  pre();
  FFT::template forward<1,LaneLayout,true,true>(x,active ? tau : 0,
                                                tb.tw[which],buf,0,lay,active,
                                                &pp,ppStride);
  if(active && colsok) {
    const long long row0=P.omBase ? sb.off_all/P.S : 0;
#pragma unroll
    for(int e=0; e < 8; ++e) {
      int l=FFT::rev(8*tau+e);
      if(l < (int) sb.nout) {
        long long a=Fbase+P.S*l+col0+lane;
        double2 v=x[0][e];
        if(P.oen) v=fmul(v,ozeta(P,(P.on*l+k0)*(col0+lane)));
        if(KIND == FFTWPP_KIND_HERMITIAN)
          ((double *) F)[a]=v.x;
        else {
          if(sb.flags & FFTWPP_SB_CONJ_OUT) v.y=-v.y;
          if(P.omBase) {
            // fused exchange: the row goes straight to its owner's buffer
            double2 *dst=(double2 *) P.omBase[row0+l]+
              (plane+P.omPlane0)*P.omStride[row0+l]+col0+lane;
            *dst=v;
          } else
            ((double2 *) F)[a]=v;
        }
      }
    }
  }
}

// r2c blocks of fftPadReal with two adjacent real columns packed into one
// complex transform z = x_a + i x_b (reference: rcfftm, convolve.cc:5794,
// 5881): after the FFT the two spectra are separated with
//   X_a[l] = (Z[l] + conj Z[M-l])/2,   X_b[l] = (Z[l] - conj Z[M-l])/(2i),
// the partner Z[M-l] being fetched through shared memory.  T/2 complex lanes.
template<int LG>
__device__ __forceinline__ void forwardSubPaired(const PlanDev& P,
                                                 const SubBlockDev& sb,
                                                 const ManyTables& tb,
                                                 const double *in,
                                                 double2 *buf, void *F,
                                                 long long Fbase, int T,
                                                 int col0, long long plane,
                                                 long long rs, int& pp,
                                                 int ppStride)
{
  typedef RegFFT<LG> FFT;
  const int M=FFT::N;
  const int TPT=FFT::TPT;
  const int TL=T/2;
  const int cl=threadIdx.x % TL;
  const int tau=threadIdx.x/TL;
  const bool active=tau < TPT;
  LaneLayout lay;
  lay.T=TL;
  lay.lane=cl;
  double2 x[1][8];
#pragma unroll
  for(int t=0; t < 8; ++t) x[0][t]=make_double2(0.0,0.0);
  if(active && (rs == T || col0+2*cl < P.C)) {
    // C is even in paired mode, so a pair is inside or outside as a whole
#pragma unroll
    for(int t=0; t < 8; ++t) {
      int s=tau+TPT*t;
      double2 acc=make_double2(0.0,0.0);
      for(int j=s; j < P.jmax; j += M) {
        double2 v=((const double2 *) in)[(j*rs)/2+cl];
        acc.x += v.x;
        acc.y += v.y;
      }
      x[0][t]=acc;
    }
  }
  FFT::template forward<1,LaneLayout,true,true>(x,active ? tau : 0,tb.tw[0],
                                                buf,0,lay,active,&pp,ppStride);
  // natural-order copy for the partner lookup
  __syncthreads();
  if(active) {
#pragma unroll
    for(int e=0; e < 8; ++e) {
      int l=FFT::rev(8*tau+e);
      buf[(l+(l >> 3))*TL+cl]=x[0][e];
    }
  }
  __syncthreads();
  if(active) {
    const bool oka=col0+2*cl < P.C;
    const bool okb=col0+2*cl+1 < P.C;
    const long long row0=P.omBase ? sb.off_all/P.S : 0;
#pragma unroll
    for(int e=0; e < 8; ++e) {
      int l=FFT::rev(8*tau+e);
      if(l < (int) sb.nout) {
        double2 z=x[0][e];
        const int lp=(M-l) & (M-1);
        double2 zp=buf[(lp+(lp >> 3))*TL+cl];
        // stored with the r2c (sign -1) convention: conj of the + transform
        double2 xa=make_double2(0.5*(z.x+zp.x),-0.5*(z.y-zp.y));
        double2 xb=make_double2(0.5*(z.y+zp.y),0.5*(z.x-zp.x));
        double2 *dst;
        if(P.omBase)
          dst=(double2 *) P.omBase[row0+l]+
            (plane+P.omPlane0)*P.omStride[row0+l]+col0+2*cl;
        else
          dst=(double2 *) F+Fbase+P.S*l+col0+2*cl;
        if(oka) dst[0]=xa;
        if(okb) dst[1]=xb;
      }
    }
  }
  __syncthreads(); // partner reads done before the buffer is reused
}

template<int LG>
__device__ __forceinline__ void backwardSubPaired(const PlanDev& P,
                                                  const SubBlockDev& sb,
                                                  const ManyTables& tb,
                                                  double *acc, double2 *buf,
                                                  const void *F,
                                                  long long Fbase, int T,
                                                  int col0, int& pp,
                                                  int ppStride)
{
  typedef RegFFT<LG> FFT;
  const int M=FFT::N;
  const int TPT=FFT::TPT;
  const int TL=T/2;
  const int cl=threadIdx.x % TL;
  const int tau=threadIdx.x/TL;
  const bool active=tau < TPT;
  LaneLayout lay;
  lay.T=TL;
  lay.lane=cl;
  double2 x[1][8];
#pragma unroll
  for(int t=0; t < 8; ++t) x[0][t]=make_double2(0.0,0.0);
  if(active) {
    const bool oka=col0+2*cl < P.C;
    const bool okb=col0+2*cl+1 < P.C;
    const double2 *src=(const double2 *) F+Fbase+col0+2*cl;
#pragma unroll
    for(int e=0; e < 8; ++e) {
      int l=FFT::rev(8*tau+e);
      // G[l] of the + transform: conj(stored[l]) for l < nout, else stored[M-l]
      bool lower=l < (int) sb.nout;
      const double2 *q=src+P.S*(lower ? l : M-l);
      double2 ga=oka ? q[0] : make_double2(0.0,0.0);
      double2 gb=okb ? q[1] : make_double2(0.0,0.0);
      if(lower) {ga.y=-ga.y; gb.y=-gb.y;}
      x[0][e]=make_double2(ga.x-gb.y,ga.y+gb.x); // ga + i gb
    }
  }
  FFT::template adjoint<1,LaneLayout,true,true>(x,active ? tau : 0,tb.tw[0],
                                                buf,0,lay,active,&pp,ppStride);
  if(active) {
#pragma unroll
    for(int t=0; t < 8; ++t) {
      int s=tau+TPT*t;
      for(int j=s; j < P.jmax; j += M) {
        double2 *a=(double2 *) acc+(j*T)/2+cl;
        double2 v=*a;
        v.x += x[0][t].x;
        v.y += x[0][t].y;
        *a=v;
      }
    }
  }
  __syncthreads();
}

// copy the twiddle tables of a plan into shared memory; returns the first
// free double2 slot
template<int LG>
__device__ __forceinline__ double2 *loadTables(const PlanDev& P,
                                               const SubBlockDev *sbs,
                                               int nsb, double2 *sm, int zlen,
                                               bool mixed, ManyTables& tb)
{
  const int M=1 << LG;
  double2 *p=sm;
  for(int w=0; w < 2; ++w) {
    tb.tw[w]=p;
    if(w == 1 && !mixed) break;
    const int n=w == 0 ? RegFFT<LG>::twCount() : RegFFT<LG-1>::twCount();
    const int want=w == 0 ? M : M/2;
    const double2 *src=NULL;
    for(int k=0; k < 2; ++k)
      if(P.tab[k].n == want) src=P.tab[k].tw8;
    if(src)
      for(int i=threadIdx.x; i < n; i += blockDim.x) p[i]=__ldg(src+i);
    p += n;
  }
  tb.zs=p;
  tb.zlen=zlen;
  if(zlen) {
    int slot=0;
    for(int isb=0; isb < nsb; ++isb) {
      const long long k0=sbs[isb].k0;
      if(k0 == 0) continue;
      for(int j=threadIdx.x; j < zlen; j += blockDim.x)
        p[(size_t) slot*zlen+j]=zeta(P,modN(P,k0,j+P.jmin));
      ++slot;
    }
    p += (size_t) slot*zlen;
  }
  return p;
}

// NT: launch bound.  The staged/gathering variants need fewer registers than
// the DIRECT ones, so their 256-thread configurations are compiled for three
// CTAs per SM (24 warps instead of 16 to hide the load latency).
template<int KIND, int LG, bool DIRECT, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 3 : 1)
fast_forward_many(PlanDev P, const SubBlockDev *__restrict__ sbs, int nsb,
                  int layout, const void *f, void *F, long long nrows,
                  long long frs, long long Frs, int T, int ntc, size_t inbytes,
                  int zlen, int mixed, long long ntiles, int pair,
                  int ppStride)
{
  typedef typename Word<KIND>::type word;
  extern __shared__ __align__(16) double2 sm2[];
  const int M=1 << LG;
  int pp=0;
  ManyTables tb;
  double2 *rest=loadTables<LG>(P,sbs,nsb,sm2,zlen,mixed != 0,tb);
  word *in=(word *) rest;
  double2 *buf=(double2 *) ((unsigned char *) rest+inbytes);
  const int lane=threadIdx.x % T;
  const int tau=threadIdx.x/T;
  __syncthreads();

  // DIRECT: the thread's 8 input points, reused by every sub-block.  (Loading
  // the next tile's points during the last sub-block's FFT was measured
  // slower: the persistent CTAs already overlap through 2 CTAs/SM.)
  double2 xin[8];
  auto loadDirect=[&](long long t_) {
    const int c0=(int) (t_ % ntc)*T;
    const double2 *gg=(const double2 *) f+(t_/ntc)*frs+c0+lane;
    const bool ok=c0+lane < P.C;
    const int TPT=M/8;
#pragma unroll
    for(int t=0; t < 8; ++t) {
      int j=tau+TPT*t;
      xin[t]=(ok && j < P.jmax) ? gg[P.S*j] : make_double2(0.0,0.0);
    }
  };

  for(long long tile=blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row=tile/ntc;
    const int col0=(int) (tile % ntc)*T;
    const bool colsok=col0+lane < P.C;
    const word *g=(const word *) f+row*frs+col0;
    {
      // pull the next tile's input rows into L2 ahead of use
      const long long nt=tile+gridDim.x;
      if(nt < ntiles) {
        const word *ng=(const word *) f+(nt/ntc)*frs+(int) (nt % ntc)*T;
        for(int j=threadIdx.x; j < P.Lin; j += blockDim.x)
          asm volatile("prefetch.global.L2 [%0];" :: "l"(ng+P.S*j));
      }
    }
    if(DIRECT) {
      loadDirect(tile);
    } else if(inbytes) {
      // stage the input tile: T contiguous words per logical row
      const int total=P.Lin*T;
      for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
        int j=idx/T;
        int c=idx-j*T;
        in[idx]=(col0+c < P.C) ? g[P.S*j+c] : word();
      }
      __syncthreads();
    }
    // inbytes == 0: every sub-block gathers straight from global memory
    const void *src=inbytes ? (const void *) in : (const void *) g;
    const long long rs=inbytes ? T : P.S;
    int slot=0;
    for(int isb=0; isb < nsb; ++isb) {
      const SubBlockDev sb=sbs[isb];
      const long long Fbase=row*Frs+(layout ? sb.off_all : sb.off_call);
      const int myslot=slot;
      if(sb.k0 != 0) ++slot;
      if(KIND == FFTWPP_KIND_REAL && !DIRECT && pair && (int) sb.mlen == M)
        forwardSubPaired<LG>(P,sb,tb,(const double *) src,buf,F,Fbase,T,col0,
                             row,rs,pp,ppStride);
      else if((int) sb.mlen == M)
        forwardSub<KIND,LG,DIRECT>(P,sb,myslot,tb,0,src,xin,buf,F,Fbase,T,
                                   col0,colsok,row,rs,pp,ppStride,
                                   NoPrefetch());
      else
        forwardSub<KIND,LG-1,false>(P,sb,myslot,tb,1,src,xin,buf,F,Fbase,T,
                                    col0,colsok,row,rs,pp,ppStride,
                                    NoPrefetch());
    }
    if(!DIRECT && inbytes) __syncthreads(); // tile reads done before restaging
  }
}

// Load the 8 scrambled-order points of one transformed sub-block owned by
// this thread, rebuilding the implied half of r2c (CONJ_OUT) blocks.
template<int KIND, int LG>
__device__ __forceinline__ void loadSpectrum(const PlanDev& P,
                                             const SubBlockDev& sb,
                                             const void *F, long long Fbase,
                                             int T, int col0, bool colsok,
                                             double2 (&x)[8])
{
  typedef RegFFT<LG> FFT;
  const int mlen=FFT::N;
  const int lane=threadIdx.x % T;
  const int tau=threadIdx.x/T;
  const bool active=tau < FFT::TPT;
#pragma unroll
  for(int e=0; e < 8; ++e) {
    double2 v=make_double2(0.0,0.0);
    if(active && colsok) {
      int l=FFT::rev(8*tau+e);
      long long a=Fbase+col0+lane;
      if(KIND == FFTWPP_KIND_HERMITIAN)
        v.x=((const double *) F)[a+P.S*l];
      else if(sb.flags & FFTWPP_SB_CONJ_OUT) {
        if(l < (int) sb.nout) {
          v=((const double2 *) F)[a+P.S*l];
          v.y=-v.y;
        } else
          v=((const double2 *) F)[a+P.S*(mlen-l)];
      } else
        v=((const double2 *) F)[a+P.S*l];
      if(P.oen) v=fmulc(v,ozeta(P,(P.on*l+sb.k0)*(col0+lane)));
    }
    x[e]=v;
  }
}

template<int KIND, int LG, bool DIRECT>
__device__ __forceinline__ void backwardSub(const PlanDev& P,
                                            const SubBlockDev& sb, int slot,
                                            const ManyTables& tb, int which,
                                            void *acc, double2 (&racc)[8],
                                            double2 *buf,
                                            const double2 (&xin)[8], int T,
                                            int& pp, int ppStride)
{
  typedef RegFFT<LG> FFT;
  const int mlen=FFT::N;
  const int TPT=FFT::TPT;
  const int lane=threadIdx.x % T;
  const int tau=threadIdx.x/T;
  const bool active=tau < TPT;
  LaneLayout lay;
  lay.T=T;
  lay.lane=lane;
  const long long k0=sb.k0;
  double2 x[1][8];
#pragma unroll
  for(int t=0; t < 8; ++t) x[0][t]=xin[t];
  FFT::template adjoint<1,LaneLayout,true,true>(x,active ? tau : 0,
                                                tb.tw[which],buf,0,lay,active,
                                                &pp,ppStride);
  if(active) {
    if(DIRECT) {
#pragma unroll
      for(int t=0; t < 8; ++t) {
        int j=tau+TPT*t;
        double2 v=x[0][t];
        if(k0 != 0 && j < P.jmax) v=fmulc(v,zetaAt(P,tb,slot,k0,j));
        racc[t]=racc[t]+v;
      }
    } else {
      const int lo=(KIND == FFTWPP_KIND_HERMITIAN) ? 0 : P.jmin;
      const int shift=(KIND == FFTWPP_KIND_CENTERED) ? P.jmin : 0;
#pragma unroll
      for(int t=0; t < 8; ++t) {
        int s=tau+TPT*t;
        int d=(s-lo) & (mlen-1);
        for(int j=lo+d; j < P.jmax; j += mlen) {
          double2 v=x[0][t];
          if(k0 != 0) v=fmulc(v,zetaAt(P,tb,slot,k0,j));
          int idx=(j-shift)*T+lane;
          if(KIND == FFTWPP_KIND_REAL) {
            double *a=(double *) acc;
            a[idx] += (sb.flags & (FFTWPP_SB_CONJ_OUT | FFTWPP_SB_SELFCONJ))
          ? v.x : 2.0*v.x;
          } else {
            double2 *a=(double2 *) acc;
            a[idx]=a[idx]+v;
          }
        }
      }
    }
  }
  if(!DIRECT) __syncthreads();
}

template<int KIND, int LG, bool DIRECT>
__global__ void __launch_bounds__(512)
fast_backward_many(PlanDev P, const SubBlockDev *__restrict__ sbs, int nsb,
                   int layout, const void *F, void *f, int accum, double scale,
                   long long nrows, long long Frs, long long frs, int T,
                   int ntc, size_t accbytes, int zlen, int mixed,
                   long long ntiles, int pair, int ppStride)
{
  typedef typename Word<KIND>::type word;
  extern __shared__ __align__(16) double2 sm2[];
  const int M=1 << LG;
  int pp=0;
  ManyTables tb;
  double2 *rest=loadTables<LG>(P,sbs,nsb,sm2,zlen,mixed != 0,tb);
  word *acc=(word *) rest;
  double2 *buf=(double2 *) ((unsigned char *) rest+accbytes);
  const int lane=threadIdx.x % T;
  const int tau=threadIdx.x/T;
  __syncthreads();

  double2 xc[8];
  for(long long tile=blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row=tile/ntc;
    const int col0=(int) (tile % ntc)*T;
    const bool colsok=col0+lane < P.C;
    word *g=(word *) f+row*frs+col0;
    const int total=P.Lin*T;
    {
      const long long nt=tile+gridDim.x;
      if(nt < ntiles && KIND != FFTWPP_KIND_HERMITIAN) {
        const long long nrow=nt/ntc;
        const int ncol=(int) (nt % ntc)*T;
        for(int isb=0; isb < nsb; ++isb) {
          const long long nb=nrow*Frs+(layout ? sbs[isb].off_all
                                       : sbs[isb].off_call)+ncol;
          const int nout=(int) sbs[isb].nout;
          for(int l=threadIdx.x; l < nout; l += blockDim.x)
            asm volatile("prefetch.global.L2 [%0];" ::
                         "l"((const double2 *) F+nb+P.S*l));
        }
      }
    }
    double2 racc[8];
#pragma unroll
    for(int t=0; t < 8; ++t) racc[t]=make_double2(0.0,0.0);
    if(DIRECT) {
      if(accum) {
        const int TPT=M/8;
#pragma unroll
        for(int t=0; t < 8; ++t) {
          int j=tau+TPT*t;
          if(colsok && j < P.jmax) racc[t]=((const double2 *) g)[P.S*j+lane];
        }
      }
    } else {
      for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
        int j=idx/T;
        int c=idx-j*T;
        acc[idx]=(accum && col0+c < P.C) ? g[P.S*j+c] : word();
      }
      __syncthreads();
    }
    int slot=0;
    if(DIRECT) {
      // software pipeline: the next sub-block's loads are in flight while
      // the current one is transformed
      const SubBlockDev s0=sbs[0];
      loadSpectrum<KIND,LG>(P,s0,F,row*Frs+(layout ? s0.off_all : s0.off_call),
                            T,col0,colsok,xc);
    }
    for(int isb=0; isb < nsb; ++isb) {
      const SubBlockDev sb=sbs[isb];
      const long long Fbase=row*Frs+(layout ? sb.off_all : sb.off_call);
      const int myslot=slot;
      if(sb.k0 != 0) ++slot;
      if(DIRECT) {
        double2 xn[8];
        if(isb+1 < nsb) {
          const SubBlockDev s1=sbs[isb+1];
          loadSpectrum<KIND,LG>(P,s1,F,
                                row*Frs+(layout ? s1.off_all : s1.off_call),
                                T,col0,colsok,xn);
        }
        backwardSub<KIND,LG,true>(P,sb,myslot,tb,0,acc,racc,buf,xc,T,pp,
                                  ppStride);
#pragma unroll
        for(int t=0; t < 8; ++t) xc[t]=xn[t];
      } else if(KIND == FFTWPP_KIND_REAL && pair && (int) sb.mlen == M) {
        backwardSubPaired<LG>(P,sb,tb,(double *) acc,buf,F,Fbase,T,col0,pp,
                              ppStride);
      } else if((int) sb.mlen == M) {
        loadSpectrum<KIND,LG>(P,sb,F,Fbase,T,col0,colsok,xc);
        backwardSub<KIND,LG,false>(P,sb,myslot,tb,0,acc,racc,buf,xc,T,pp,
                                   ppStride);
      } else {
        loadSpectrum<KIND,LG-1>(P,sb,F,Fbase,T,col0,colsok,xc);
        backwardSub<KIND,LG-1,false>(P,sb,myslot,tb,1,acc,racc,buf,xc,T,pp,
                                     ppStride);
      }
    }
    if(DIRECT) {
      const int TPT=M/8;
      if(colsok) {
#pragma unroll
        for(int t=0; t < 8; ++t) {
          int j=tau+TPT*t;
          if(j < P.jmax) {
            if(P.omBase)
              ((double2 *) P.omBase[j])[(row+P.omPlane0)*P.omStride[j]+col0+lane]=
                wscale(racc[t],scale);
            else
              ((double2 *) g)[P.S*j+lane]=wscale(racc[t],scale);
          }
        }
      }
    } else {
      for(int idx=threadIdx.x; idx < total; idx += blockDim.x) {
        int j=idx/T;
        int c=idx-j*T;
        if(col0+c < P.C)
          g[P.S*j+c]=wscale(acc[idx],scale);
      }
      __syncthreads(); // tile stored before the next tile re-initialises it
    }
  }
}

// ---------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------

const size_t SMEM_MAX=227*1024;

template<class K>
int allowSmem(K kernel)
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

int ilog2(unsigned v)
{
  int l=0;
  while((1u << l) < v) ++l;
  return l;
}

bool ispow2(unsigned v) {return v && !(v & (v-1));}

size_t wordBytes(int kind)
{
  return kind == FFTWPP_KIND_REAL ? sizeof(double) : sizeof(double2);
}

// Tuning constants.  They were found with A/B runs on B200 (profiles/README.md)
// through environment switches; those switches exist only in builds with
// -DFFTWPP_EXPERIMENT_SWITCHES, the product reads no tuning knobs at run time
// (FFTWPP_NO_FAST, which forces the generic kernels, is kept for testing).
#ifdef FFTWPP_EXPERIMENT_SWITCHES
static int envInt(const char *name, int def)
{
  const char *s=getenv(name);
  return (s && *s) ? atoi(s) : def;
}
#define TUNE(name, def) ([]{static int v=envInt(name,def); return v;}())
#else
#define TUNE(name, def) (def)
#endif

// complex lanes per strided tile (64-byte global runs; measured: 2 CTAs/SM of
// 256 threads beat one CTA with 8 lanes, 17.7 vs 19.0 ms at 512^3)
int tileLanes()
{
  int T=TUNE("FFTWPP_TILE_LANES",4);
  if(T != 2 && T != 4 && T != 8 && T != 16) T=4;
  return T;
}

// ping-pong exchange buffers: measured no gain (the barriers are not the limiter)
bool pingpongEnabled() {return TUNE("FFTWPP_PINGPONG",0) != 0;}

int realLanes()
{
  int T=TUNE("FFTWPP_TILE_LANES_REAL",2*tileLanes());
  if(T != 4 && T != 8 && T != 16 && T != 32) T=2*tileLanes();
  return T;
}

// stage the real input tile in shared memory instead of gathering it
bool stageDisabled() {return TUNE("FFTWPP_STAGE_REAL",0) != 0;}

// r2c column pairing off
bool pairDisabled() {return TUNE("FFTWPP_NO_PAIR",0) != 0;}

// software-pipelined fused convolution (0: the plain kernel)
bool convPipeEnabled() {return TUNE("FFTWPP_CONV_PIPE",1) != 0;}

// p=1, q=2 rows on the specialised straight-line kernel (0: general kernel)
bool convQ2Disabled() {return TUNE("FFTWPP_CONV_Q2",1) == 0;}

// three CTAs per SM for the gathering real forward pass (+15 %)
bool threeCtasEnabled() {return TUNE("FFTWPP_THREE_CTAS",1) != 0;}

bool fastDisabled()
{
  static int off=-1;
  if(off < 0) {
    const char *s=getenv("FFTWPP_NO_FAST");
    off=(s && *s && *s != '0') ? 1 : 0;
  }
  return off == 1;
}

#define LG_CASES(CALL)            \
  switch(lg) {                    \
    case 4: {CALL(4); break;}     \
    case 5: {CALL(5); break;}     \
    case 6: {CALL(6); break;}     \
    case 7: {CALL(7); break;}     \
    case 8: {CALL(8); break;}     \
    case 9: {CALL(9); break;}     \
    case 10: {CALL(10); break;}   \
    case 11: {CALL(11); break;}   \
    case 12: {CALL(12); break;}   \
    default: return 0;            \
  }

struct ManyGeom {
  int T, nthreads, ntc, zlen, mixed, pair, ppStride;
  size_t tilebytes, smem;
  uint64_t ntiles, grid;
  bool direct;
};

// Tile geometry, shared-memory budget and persistent grid of a Many pass.
template<int KIND>
int manyGeometry(Plan *pl, int lg, uint64_t nrows, ManyGeom& g,
                 bool forward=false)
{
  FastInfo *fi=pl->fast;
  const int M=1 << lg;
  g.mixed=fi->uniform ? 0 : 1;
  g.direct=fi->uniform && KIND == FFTWPP_KIND_COMPLEX && fi->nterm == 1;

  int twn=0;
  for(int k=0; k < lg/3; ++k)
    if(lg-3*(k+1) > 0) twn += 7 << (lg-3*(k+1));
  if(g.mixed)
    for(int k=0; k < (lg-1)/3; ++k)
      if(lg-1-3*(k+1) > 0) twn += 7 << (lg-1-3*(k+1));
  int nz=0;
  for(size_t i=0; i < pl->hsub.size(); ++i) nz += pl->hsub[i].k0 != 0;
  int span=pl->dev.jmax-pl->dev.jmin;
  int T=tileLanes();
  g.pair=(KIND == FFTWPP_KIND_REAL && fi->pairable && pl->dev.C >= 2 &&
          pl->dev.C % 2 == 0 && pl->dev.S % 2 == 0 && !pairDisabled()) ? 1 : 0;
  if(g.pair) T=realLanes(); // 8-byte words: more lanes for the same bytes
  while(T*(M/8) < 256 && T < 128) T *= 2; // short transforms: wider tiles
  while(T > 1 && (size_t) T > pl->dev.C) T /= 2;
  if(g.pair && T < 2) g.pair=0;
  const int tdiv=g.pair ? 2 : 1; // threads per lane: M/8, or M/16 when paired
  for(;;) {
    // forward passes of paired real plans gather straight from global memory
    const bool notile=g.direct || (forward && g.pair && !stageDisabled());
    g.tilebytes=notile ? 0 :
      (((size_t) pl->dev.Lin*T*wordBytes(KIND)+15) & ~(size_t) 15);
    // exchange buffer: M points per lane; paired r2c blocks use T/2 complex
    // lanes plus one padding row per 8 for the natural-order partner lookup
    size_t bufwords=g.pair ? (size_t) (M+M/8)*(T/2) : (size_t) M*T;
    // ping-pong exchange buffers (one barrier per exchange) when two CTAs
    // still fit per SM, else a single buffer with two barriers
    size_t fixed=(size_t) twn*sizeof(double2)+g.tilebytes+
      (size_t) std::min<size_t>(nz*span*sizeof(double2),48*1024);
    bool pingpong=pingpongEnabled() &&
      fixed+2*bufwords*sizeof(double2) <= 112*1024;
    g.ppStride=pingpong ? (int) bufwords : 0;
    size_t base=(size_t) twn*sizeof(double2)+g.tilebytes+
      (pingpong ? 2 : 1)*bufwords*sizeof(double2);
    size_t zbytes=(size_t) nz*span*sizeof(double2);
    g.zlen=span;
    if(zbytes > 48*1024 || base+zbytes > SMEM_MAX) {g.zlen=0; zbytes=0;}
    g.smem=base+zbytes;
    if(g.smem <= SMEM_MAX && T*(M/8)/tdiv <= 512) break;
    if(T == 1 || (g.pair && T == 2)) return 0;
    T /= 2;
  }
  g.T=T;
  g.nthreads=T*(M/8)/tdiv;
  if(g.nthreads < 32) return 0;
  g.ntc=(int) ((pl->dev.C+T-1)/T);
  g.ntiles=nrows*(uint64_t) g.ntc;
  // persistent grid: as many CTAs as fit on the device (by shared memory and
  // threads), a few waves for load balance
  size_t perSM=std::min<size_t>(std::max<size_t>(1,(227*1024)/(g.smem+1024)),
                                2048/g.nthreads);
  g.grid=std::min<uint64_t>(g.ntiles,(uint64_t) sm_count()*perSM*4);
  return 1;
}

template<int KIND>
int launchForwardMany(Plan *pl, int lg, uint64_t sb0, uint64_t nsb, int layout,
                      const void *f, void *F, uint64_t nrows, uint64_t frs,
                      uint64_t Frs, cudaStream_t st,
                      const unsigned long long *omBase,
                      const long long *omStride, long long omPlane0)
{
  ManyGeom g;
  if(!manyGeometry<KIND>(pl,lg,nrows,g,true)) return 0;
  if(g.ntiles == 0) return 1;
  // paired real columns are read as 16-byte words
  if(g.pair && ((frs & 1) || ((uintptr_t) f & 15))) return 0;
  PlanDev dev=pl->dev;
  dev.omBase=omBase;
  dev.omStride=omStride;
  dev.omPlane0=omPlane0;
  // zeta rows are indexed by sub-block slot within [sb0,sb0+nsb)
  int rc=0;
#define CALLD(LGV, DIR, NTV)                                                 \
  rc=allowSmem(fast_forward_many<KIND,LGV,DIR,NTV>);                         \
  if(rc) return rc;                                                          \
  prof_begin(4*pl->tag+0,st);                                                \
  fast_forward_many<KIND,LGV,DIR,NTV>                                        \
    <<<(unsigned) g.grid,g.nthreads,g.smem,st>>>                             \
    (dev,pl->dsub+sb0,(int) nsb,layout,f,F,(long long) nrows,            \
     (long long) frs,(long long) Frs,g.T,g.ntc,g.tilebytes,g.zlen,g.mixed,   \
     (long long) g.ntiles,g.pair,g.ppStride);
  // measured (512^3, B200): 3 CTAs/SM speeds the gathering real pass up by
  // 15 %; bounding the DIRECT complex passes the same way (spills) or letting
  // them re-gather their inputs per sub-block costs 35-40 %
  const bool three=KIND == FFTWPP_KIND_REAL && !g.direct &&
    g.nthreads <= 256 && g.tilebytes == 0 && threeCtasEnabled();
#define CALL(LGV)                                                            \
  if(g.direct) {CALLD(LGV,true,512)}                                         \
  else if(KIND == FFTWPP_KIND_REAL && three) {CALLD(LGV,false,256)}          \
  else {CALLD(LGV,false,512)}
  LG_CASES(CALL)
#undef CALL
#undef CALLD
  rc=check_launch("fast_forward_many",st);
  return rc ? rc : 1;
}

template<int KIND>
int launchBackwardMany(Plan *pl, int lg, uint64_t sb0, uint64_t nsb,
                       int layout, const void *F, void *f, int accumulate,
                       double scale, uint64_t nrows, uint64_t Frs,
                       uint64_t frs, cudaStream_t st,
                       const unsigned long long *omBase,
                       const long long *omStride, long long omPlane0)
{
  ManyGeom g;
  if(!manyGeometry<KIND>(pl,lg,nrows,g)) return 0;
  if(g.ntiles == 0) return 1;
  if(g.pair && ((frs & 1) || ((uintptr_t) f & 15))) return 0;
  if(omBase && !g.direct) return 0; // mapped output: register variant only
  PlanDev dev=pl->dev;
  dev.omBase=omBase;
  dev.omStride=omStride;
  dev.omPlane0=omPlane0;
  int rc=0;
#define CALLD(LGV, DIR)                                                      \
  rc=allowSmem(fast_backward_many<KIND,LGV,DIR>);                            \
  if(rc) return rc;                                                          \
  prof_begin(4*pl->tag+1,st);                                                \
  fast_backward_many<KIND,LGV,DIR><<<(unsigned) g.grid,g.nthreads,g.smem,st>>> \
    (dev,pl->dsub+sb0,(int) nsb,layout,F,f,accumulate,scale,             \
     (long long) nrows,(long long) Frs,(long long) frs,g.T,g.ntc,            \
     g.tilebytes,g.zlen,g.mixed,(long long) g.ntiles,g.pair,g.ppStride);
#define CALL(LGV) if(g.direct) {CALLD(LGV,true)} else {CALLD(LGV,false)}
  LG_CASES(CALL)
#undef CALL
#undef CALLD
  rc=check_launch("fast_backward_many",st);
  return rc ? rc : 1;
}

template<int NTERM>
int launchConvRows(Plan *pl, int lg, void *const *f, int mult, double scale,
                   uint64_t nrows, uint64_t rs, cudaStream_t st)
{
  const int M=1 << lg;
  const int TPT=M/8;
  const int NT=TPT > 256 ? TPT : 256;
  const int ROWS=NT/TPT;
  const int BUF=NTERM*M+NTERM*M/8;
  int nr8=lg/3, twn=0;
  for(int k=0; k < nr8; ++k)
    if(lg-3*(k+1) > 0) twn += 7 << (lg-3*(k+1));
  int nz=0;
  for(size_t i=0; i < pl->hsub.size(); ++i) nz += pl->hsub[i].k0 != 0;
  int zlen=std::min<int>(pl->dev.jmax,NTERM*M);
  size_t base=(size_t) twn*sizeof(double2)+2*(size_t) ROWS*BUF*sizeof(double2);
  size_t zbytes=(size_t) nz*zlen*sizeof(double2);
  const size_t budget=NT > 256 ? SMEM_MAX : 113*1024;
  if(base+zbytes > budget) {zlen=0; zbytes=0;} // residue twiddles from L1/L2
  size_t smem=base+zbytes;
  if(smem > SMEM_MAX) return 0;
  uint64_t ngroups=(nrows+ROWS-1)/ROWS;
  if(ngroups == 0) return 1;
  uint64_t grid=std::min<uint64_t>(ngroups,(uint64_t) sm_count()*2*4);
  int tabid=0;
  int rc=0;
  // p=1, q=2, full rows: the straight-line kernel with register twiddles
  if(NTERM == 1 && convPipeEnabled() && !convQ2Disabled() &&
     pl->hsub.size() == 2 && pl->hsub[0].k0 == 0 && pl->hsub[1].k0 != 0 &&
     pl->dev.jmax == M && pl->dev.jmin == 0 &&
     (mult == FFTWPP_MULT_BINARY || mult == FFTWPP_MULT_CORRELATION)) {
    size_t sm2=((size_t) twn+TPT+2*(size_t) ROWS*BUF)*sizeof(double2);
    if(sm2 > SMEM_MAX) return 0;
    // zeta_N^{k0 TPT}, the step between a thread's successive points
    const long double ang=2.0L*3.141592653589793238462643383279502884L*
      (long double) ((pl->hsub[1].k0*(unsigned long long) TPT) %
                     (unsigned long long) pl->dev.N)/(long double) pl->dev.N;
    const double2 zstep=make_double2((double) cosl(ang),(double) sinl(ang));
#define CALLQ(LGV, MU)                                                       \
    rc=allowSmem(fast_conv_rows_q2<LGV,MU>);                                 \
    if(rc) return rc;                                                        \
    prof_begin(4*pl->tag+2,st);                                              \
    fast_conv_rows_q2<LGV,MU><<<(unsigned) grid,NT,sm2,st>>>                 \
      (pl->dev,pl->dsub,(double2 *) f[0],(const double2 *) f[1],scale,       \
       zstep,(long long) nrows,(long long) rs,tabid,(long long) ngroups);
#define CALL(LGV)                                                            \
    if(mult == FFTWPP_MULT_BINARY) {CALLQ(LGV,FFTWPP_MULT_BINARY)}           \
    else {CALLQ(LGV,FFTWPP_MULT_CORRELATION)}
    LG_CASES(CALL)
#undef CALL
#undef CALLQ
    rc=check_launch("fast_conv_rows_q2",st);
    return rc ? rc : 1;
  }
  if(NTERM == 1 && convPipeEnabled()) {
#define CALL(LGV)                                                            \
    rc=allowSmem(fast_conv_rows_pipe<LGV>);                                  \
    if(rc) return rc;                                                        \
    prof_begin(4*pl->tag+2,st);                                              \
    fast_conv_rows_pipe<LGV><<<(unsigned) grid,NT,smem,st>>>                 \
      (pl->dev,pl->dsub,(int) pl->hsub.size(),(double2 *) f[0],              \
       (const double2 *) f[1],mult,scale,(long long) nrows,(long long) rs,   \
       tabid,zlen,(long long) ngroups);
    LG_CASES(CALL)
#undef CALL
    rc=check_launch("fast_conv_rows_pipe",st);
    return rc ? rc : 1;
  }
#define CALL(LGV)                                                            \
  rc=allowSmem(fast_conv_rows<LGV,NTERM>);                                   \
  if(rc) return rc;                                                          \
  prof_begin(4*pl->tag+2,st);                                                \
  fast_conv_rows<LGV,NTERM><<<(unsigned) grid,NT,smem,st>>>                  \
    (pl->dev,pl->dsub,(int) pl->hsub.size(),(double2 *) f[0],                \
     (const double2 *) f[1],mult,scale,(long long) nrows,(long long) rs,     \
     tabid,zlen,(long long) ngroups);
  LG_CASES(CALL)
#undef CALL
  rc=check_launch("fast_conv_rows",st);
  return rc ? rc : 1;
}

int launchConvRowsHerm(Plan *pl, int lg, void *const *f, double scale,
                       uint64_t nrows, uint64_t rs, cudaStream_t st)
{
  const int M=1 << lg;
  const int TPT=M/8;
  const int NT=TPT > 256 ? TPT : 256;
  const int ROWS=NT/TPT;
  const int BUF=M+M/8;
  int nr8=lg/3, twn=0;
  for(int k=0; k < nr8; ++k)
    if(lg-3*(k+1) > 0) twn += 7 << (lg-3*(k+1));
  int nz=0;
  for(size_t i=0; i < pl->hsub.size(); ++i) nz += pl->hsub[i].k0 != 0;
  int zlen=pl->dev.jmax-pl->dev.jmin;
  size_t base=(size_t) twn*sizeof(double2)+2*(size_t) ROWS*BUF*sizeof(double2);
  size_t zbytes=(size_t) nz*zlen*sizeof(double2);
  const size_t budget=NT > 256 ? SMEM_MAX : 113*1024;
  if(base+zbytes > budget) {zlen=0; zbytes=0;}
  size_t smem=base+zbytes;
  if(smem > SMEM_MAX) return 0;
  uint64_t ngroups=(nrows+ROWS-1)/ROWS;
  if(ngroups == 0) return 1;
  uint64_t grid=std::min<uint64_t>(ngroups,(uint64_t) sm_count()*2*4);
  int tabid=0;
  int rc=0;
#define CALL(LGV)                                                            \
  rc=allowSmem(fast_conv_rows_herm<LGV>);                                    \
  if(rc) return rc;                                                          \
  prof_begin(4*pl->tag+2,st);                                                \
  fast_conv_rows_herm<LGV><<<(unsigned) grid,NT,smem,st>>>                   \
    (pl->dev,pl->dsub,(int) pl->hsub.size(),(double2 *) f[0],                \
     (const double2 *) f[1],scale,(long long) nrows,(long long) rs,tabid,    \
     zlen,(long long) ngroups);
  LG_CASES(CALL)
#undef CALL
  rc=check_launch("fast_conv_rows_herm",st);
  return rc ? rc : 1;
}

} // namespace

void fast_plan_init(Plan *pl)
{
  pl->fast=NULL;
  if(fastDisabled()) return;
  unsigned mmax=pl->mmax;
  if(!ispow2(mmax) || mmax < 16 || mmax > 4096) return;
  bool uniform=true;
  unsigned mmin=mmax;
  for(size_t i=0; i < pl->hsub.size(); ++i) {
    unsigned ml=pl->hsub[i].mlen;
    if(ml != mmax) {
      uniform=false;
      if(2*ml != mmax) return;
      mmin=ml;
    }
  }
  if(pl->dev.Lin > 32768) return;
  FastInfo *fi=new FastInfo;
  fi->log2m=ilog2(mmax);
  fi->uniform=uniform;
  fi->pairable=pl->dev.kind == FFTWPP_KIND_REAL;
  for(size_t i=0; i < pl->hsub.size(); ++i)
    if(pl->hsub[i].mlen == mmax && !(pl->hsub[i].flags & FFTWPP_SB_CONJ_OUT))
      fi->pairable=false;
  if(!uniform) {
    // the half-length blocks then use the same thread count as the pairs
    for(size_t i=0; i < pl->hsub.size(); ++i)
      if(pl->hsub[i].mlen != mmax &&
         (pl->hsub[i].flags & FFTWPP_SB_CONJ_OUT))
        fi->pairable=false;
  }
  int span=pl->dev.jmax-pl->dev.jmin;
  fi->nterm=(span+(int) mmin-1)/(int) mmin;
  pl->fast=fi;
}

void fast_plan_free(Plan *pl)
{
  delete pl->fast;
  pl->fast=NULL;
}

int fast_try_forward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                     const void *f, void *F, uint64_t nrows, uint64_t frs,
                     uint64_t Frs, cudaStream_t st,
                     const unsigned long long *omBase,
                     const long long *omStride, long long omPlane0)
{
  FastInfo *fi=pl->fast;
  if(!fi || pl->dev.C < 2) return 0;
  int lg=fi->log2m;
  if(!omBase) { // TMA-staged tiles where a tile shape is instantiated
    int rc=tma_try_forward(pl,sb0,nsb,layout,f,F,nrows,frs,Frs,st);
    if(rc) return rc;
  }
  switch(pl->dev.kind) {
    case FFTWPP_KIND_COMPLEX:
      return launchForwardMany<FFTWPP_KIND_COMPLEX>(pl,lg,sb0,nsb,layout,f,F,
                                                    nrows,frs,Frs,st,omBase,
                                                    omStride,omPlane0);
    case FFTWPP_KIND_CENTERED:
      return launchForwardMany<FFTWPP_KIND_CENTERED>(pl,lg,sb0,nsb,layout,f,F,
                                                     nrows,frs,Frs,st,omBase,
                                                     omStride,omPlane0);
    case FFTWPP_KIND_REAL:
      return launchForwardMany<FFTWPP_KIND_REAL>(pl,lg,sb0,nsb,layout,f,F,
                                                 nrows,frs,Frs,st,omBase,
                                                 omStride,omPlane0);
  }
  return 0;
}

int fast_try_backward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                      const void *F, void *f, int accumulate, double scale,
                      uint64_t nrows, uint64_t Frs, uint64_t frs,
                      cudaStream_t st, const unsigned long long *omBase,
                      const long long *omStride, long long omPlane0)
{
  FastInfo *fi=pl->fast;
  if(!fi || pl->dev.C < 2) return 0;
  int lg=fi->log2m;
  if(!omBase) {
    int rc=tma_try_backward(pl,sb0,nsb,layout,F,f,accumulate,scale,nrows,Frs,
                            frs,st);
    if(rc) return rc;
  }
  switch(pl->dev.kind) {
    case FFTWPP_KIND_COMPLEX:
      return launchBackwardMany<FFTWPP_KIND_COMPLEX>(pl,lg,sb0,nsb,layout,F,f,
                                                     accumulate,scale,nrows,
                                                     Frs,frs,st,omBase,
                                                     omStride,omPlane0);
    case FFTWPP_KIND_CENTERED:
      return launchBackwardMany<FFTWPP_KIND_CENTERED>(pl,lg,sb0,nsb,layout,F,f,
                                                      accumulate,scale,nrows,
                                                      Frs,frs,st,omBase,
                                                      omStride,omPlane0);
    case FFTWPP_KIND_REAL:
      return launchBackwardMany<FFTWPP_KIND_REAL>(pl,lg,sb0,nsb,layout,F,f,
                                                  accumulate,scale,nrows,Frs,
                                                  frs,st,omBase,omStride,omPlane0);
  }
  return 0;
}

// Would fast_try_forward / fast_try_backward accept a mapped-output (fused
// exchange) launch of this plan?  Geometry only -- nothing is launched.
int fast_mapped_supported(Plan *pl, int backward)
{
  FastInfo *fi=pl->fast;
  if(!fi || pl->dev.C < 2) return 0;
  ManyGeom g;
  int ok=0;
  switch(pl->dev.kind) {
    case FFTWPP_KIND_COMPLEX:
      ok=manyGeometry<FFTWPP_KIND_COMPLEX>(pl,fi->log2m,1,g,!backward);
      break;
    case FFTWPP_KIND_CENTERED:
      ok=manyGeometry<FFTWPP_KIND_CENTERED>(pl,fi->log2m,1,g,!backward);
      break;
    case FFTWPP_KIND_REAL:
      ok=manyGeometry<FFTWPP_KIND_REAL>(pl,fi->log2m,1,g,!backward);
      break;
    default:
      return 0;
  }
  if(!ok) return 0;
  if(backward && !g.direct) return 0; // mapped output: register variant only
  return 1;
}

int fast_try_convolve(Plan *pl, void *const *f, uint32_t A, uint32_t B,
                      int mult, double scale, uint64_t nrows, uint64_t rs,
                      cudaStream_t st)
{
  FastInfo *fi=pl->fast;
  if(!fi || !fi->uniform) return 0;
  if(pl->dev.C != 1 || pl->dev.S != 1) return 0;
  if(A != 2 || B != 1) return 0;
  if(pl->dev.kind == FFTWPP_KIND_HERMITIAN) {
    // stored modes 0..H-1 with H <= m: every W[s] has at most the two terms
    // j=s and j=s-m
    if(mult != FFTWPP_MULT_REALBINARY) return 0;
    if(pl->dev.jmax > (1 << fi->log2m)) return 0;
    return launchConvRowsHerm(pl,fi->log2m,f,scale,nrows,rs,st);
  }
  if(pl->dev.kind != FFTWPP_KIND_COMPLEX) return 0;
  if(mult != FFTWPP_MULT_BINARY && mult != FFTWPP_MULT_CORRELATION) return 0;
  if(fi->nterm == 1)
    return launchConvRows<1>(pl,fi->log2m,f,mult,scale,nrows,rs,st);
  if(fi->nterm == 2)
    return launchConvRows<2>(pl,fi->log2m,f,mult,scale,nrows,rs,st);
  return 0;
}

} // namespace fftwpp_gpu
