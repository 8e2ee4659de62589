// tma_kernels.cu -- TMA-staged strided ("Many") residue passes for sm_100a.
//
// Same mathematics as fast_forward_many / fast_backward_many (fast_kernels.cu;
// reference fftPad::forward1Many/forward2Many + backward, convolve.cc:984-1225,
// 1573-1763): a tile of T adjacent columns is transformed along the strided
// index.  What changes is how the tile moves:
//
//   HBM --cp.async.bulk.tensor (TMA, one elected thread)--> shared memory
//        completion on an mbarrier (complete_tx::bytes); no thread issues a
//        global load, no register is live across the HBM latency, and the
//        next tile is in flight while this one is transformed;
//   register FFT (regfft.cuh), exchanges through padded shared memory;
//   results -> shared memory tile in natural row order -> ONE bulk tensor
//        store per box (cp.async.bulk.tensor.global.shared::cta), tracked
//        with bulk groups, so no thread issues a global store either.
//
// Measured motivation (profiles/README.md, round 2): the strided passes are
// bound by LSU data-pipe wavefronts, and a 128-bit global access of a 64-byte
// row costs ~7.2 wavefronts against 4 for shared memory; the real x pass is
// bound by exposed load latency.  TMA removes both.
//
// Bank conflicts with T=4 lanes: a quarter-warp (8 threads = one shared-memory
// wavefront of a 128-bit access) holds 2 consecutive taus x 4 lanes, i.e. two
// 64-byte rows; they must fall into different halves of the 128-byte bank
// window.  Exchanges: positions are padded, p -> p+(p>>3), which gives rows of
// different parity for every pass.  Staged tiles are dense (the TMA side cannot
// pad), so the digit-reversed accesses of the complex passes (rows 8 apart)
// keep a 2-way conflict; the real x pass avoids it by ordering its two 128-bit
// accesses by the parity of tau.  (128-byte swizzled tensor maps were tried
// for the complex tiles: with a 64-byte inner box they fault on B200.)

#include "regfft.cuh"

#include <cuda.h>

#include <mutex>

#ifndef FFTWPP_TMA_L2PROMO_DEFAULT
#define FFTWPP_TMA_L2PROMO_DEFAULT 0
#endif

namespace fftwpp_gpu {

namespace {

// ---------------------------------------------------------------------------
// PTX wrappers: mbarrier, bulk tensor copies, proxy fences
// ---------------------------------------------------------------------------

__device__ __forceinline__ unsigned smemAddr(const void *p)
{
  return (unsigned) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbarInit(void *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::
               "r"(smemAddr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbarInitFence()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbarExpectTx(void *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::
               "r"(smemAddr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbarTry(void *bar, unsigned parity)
{
  unsigned ok;
  asm volatile(
    "{\n\t"
    ".reg .pred p;\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
    "selp.u32 %0, 1, 0, p;\n\t"
    "}" : "=r"(ok) : "r"(smemAddr(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// A bulk copy that never completes (a wrong byte count, a bad tensor map)
// must fail the launch, not hang the device: give up after ~2 s.
__device__ __forceinline__ void mbarWait(void *bar, unsigned parity)
{
  if(mbarTry(bar,parity)) return;
  const long long t0=clock64();
  while(!mbarTry(bar,parity))
    if(clock64()-t0 > 4000000000ll) __trap();
}

// global (tensor map, coordinates) -> shared, completion on mbarrier
__device__ __forceinline__ void tmaLoad3(void *dst, const CUtensorMap *map,
                                         void *bar, int c0, int c1, int c2)
{
  asm volatile(
    "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
    " [%0], [%1, {%3, %4, %5}], [%2];" ::
    "r"(smemAddr(dst)), "l"((unsigned long long) map), "r"(smemAddr(bar)),
    "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// shared -> global (tensor map, coordinates), tracked by the bulk group
__device__ __forceinline__ void tmaStore3(const CUtensorMap *map,
                                          const void *src, int c0, int c1,
                                          int c2)
{
  asm volatile(
    "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group"
    " [%0, {%2, %3, %4}], [%1];" ::
    "l"((unsigned long long) map), "r"(smemAddr(src)), "r"(c0), "r"(c1),
    "r"(c2) : "memory");
}

__device__ __forceinline__ void tmaCommit()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// wait until at most N of this thread's bulk groups still READ shared memory
template<int N>
__device__ __forceinline__ void tmaWaitRead()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}

// make generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fenceAsyncShared()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------
// output destinations
// ---------------------------------------------------------------------------
//
// A pass writes its output rows to one or more destinations, each owning a
// contiguous range of rows: the local buffer (one destination), or -- fused
// exchange of the distributed convolutions -- the buffers of the peer GPUs
// (CUDA-IPC mapped), rows split as the slab decomposition splits them.  The
// kernels stage their output as a few fixed boxes of rows per tile ("slots":
// sub-block x box); the host cuts every slot into SEGMENTS, one per owner it
// overlaps, and encodes one tensor map per segment whose box height is
// exactly the segment's row count.  A slot is then stored with one bulk
// tensor store per segment, always inside the tensor bounds (measured: a
// store whose box starts at a negative coordinate raises an illegal-
// instruction fault on B200, so nothing relies on clipping).  The peer copies
// are issued by the TMA unit over NVLink; no thread executes a remote store.
static const int MAXSEG=24;
static const int MAXSLOT=8;

struct DestSet {
  CUtensorMap map[MAXSEG];
  int srcRow[MAXSEG];   // first row of the segment inside the slot's box
  int dstRow[MAXSEG];   // row coordinate in the owner's tensor
  int first[MAXSLOT+1]; // slot s owns segments [first[s],first[s+1])
  int plane0;
};

// src: first row of the slot's box in shared memory, rowWords double2 per row
__device__ __forceinline__ void storeSlot(const DestSet& D, int slot,
                                          const double2 *src, int rowWords,
                                          int c0, int plane)
{
  for(int i=D.first[slot]; i < D.first[slot+1]; ++i)
    tmaStore3(&D.map[i],src+D.srcRow[i]*rowWords,c0,D.dstRow[i],
              D.plane0+plane);
}

// ---------------------------------------------------------------------------
// tile geometry
// ---------------------------------------------------------------------------

template<int LG, int NTHR=256>
struct TileGeom {
  static const int M=1 << LG;
  static const int TPT=M/8;               // threads per column
  static const int NT=NTHR;               // threads per CTA
  static const int T=NT/TPT;              // columns (lanes) per tile
  static const int BR=M < 256 ? M : 256;  // rows per TMA box
  static const int NBOX=M/BR;
  static const int TILE=M*T;              // dense tile, double2 words
  static const int EPAD=(M+M/8)*T;        // padded exchange buffer
  static const int ZT=TPT+1;              // residue twiddles per slot:
                                          // zeta^{k0 tau}, then zeta^{k0 TPT}
};

// Exchange layout: element p of lane `lane` at (p+(p>>3))*T+lane.
template<int T>
struct PadLaneLayout {
  int lane;
  __device__ __forceinline__ int addr(int p) const {
    return (p+(p >> 3))*T+lane;
  }
  __device__ __forceinline__ void sync() const {__syncthreads();}
};

template<int LG, int NTHR>
__device__ __forceinline__ void threadMap(int& lane, int& tau)
{
  typedef TileGeom<LG,NTHR> G;
  lane=threadIdx.x % G::T;
  tau=threadIdx.x/G::T;
}

// double2 index of (row r, lane) inside a staged (dense) tile
template<int T>
__device__ __forceinline__ int tileAddr(int r, int lane)
{
  return r*T+lane;
}

// 1024-byte aligned start of the dynamic shared memory (swizzle atoms)
__device__ __forceinline__ unsigned char *alignedSmem(unsigned char *raw)
{
  const unsigned a=smemAddr(raw);
  return raw+((1024u-(a & 1023u)) & 1023u);
}

// residue twiddles of the thread's 8 points j=tau+TPT*t:
// zeta^{k0 j} = zeta^{k0 tau} * (zeta^{k0 TPT})^t, by successive products
template<int LG, bool CONJ>
__device__ __forceinline__ void applyZeta(double2 (&x)[8], const double2 *zt,
                                          int tau)
{
  typedef TileGeom<LG> G;
  double2 z=zt[tau];
  const double2 step=zt[G::TPT];
#pragma unroll
  for(int t=0; t < 8; ++t) {
    x[t]=CONJ ? fmulc(x[t],z) : fmul(x[t],z);
    if(t < 7) z=fmul(z,step);
  }
}

// residue twiddle tables of the sub-blocks with k0 != 0 -> shared memory
template<int LG>
__device__ __forceinline__ void loadZeta(const PlanDev& P,
                                         const SubBlockDev *sbs, int nsb,
                                         double2 *zt)
{
  typedef TileGeom<LG> G;
  int slot=0;
  for(int isb=0; isb < nsb; ++isb) {
    const long long k0=sbs[isb].k0;
    if(k0 == 0) continue;
    for(int j=threadIdx.x; j <= G::TPT; j += blockDim.x)
      zt[slot*G::ZT+j]=zeta(P,modN(P,k0,j));
    ++slot;
  }
}

// ---------------------------------------------------------------------------
// forward pass, uniform complex plans with L <= m (one term per W[s])
// ---------------------------------------------------------------------------
//
// shared memory: [input tile M*T][E0 padded][E1 padded][zeta slots][mbarrier]
// E0/E1 alternate between sub-blocks: exchange buffer first, then the
// natural-order output tile that the bulk store reads.
template<int LG>
__global__ void __launch_bounds__(256,2)
tma_forward_direct(const __grid_constant__ CUtensorMap tmIn,
                   const __grid_constant__ DestSet out, PlanDev P,
                   const SubBlockDev *__restrict__ sbs, int nsb, int layout,
                   int ntc, long long ntiles, int tabid)
{
  typedef TileGeom<LG> G;
  typedef WFFT<LG> FFT;
  extern __shared__ __align__(128) unsigned char smraw[];
  double2 *inS=(double2 *) alignedSmem(smraw);
  double2 *E0=inS+G::TILE;
  double2 *zt=inS+G::TILE+2*G::EPAD;
  int nz=0;
  for(int isb=0; isb < nsb; ++isb) nz += sbs[isb].k0 != 0;
  unsigned long long *full=(unsigned long long *) (zt+nz*G::ZT);

  int lane,tau;
  threadMap<LG,256>(lane,tau);
  PadLaneLayout<G::T> lay;
  lay.lane=lane;
  double2 w1[FFT::NR8 > 0 ? FFT::NR8 : 1];
  FFT::loadW(P.tab[tabid].tw8,tau,w1);
  loadZeta<LG>(P,sbs,nsb,zt);
  const unsigned tileBytes=G::TILE*sizeof(double2);

  auto issueLoad=[&](long long tile) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*G::T;
    mbarExpectTx(full,tileBytes);
#pragma unroll
    for(int b=0; b < G::NBOX; ++b)
      tmaLoad3(inS+b*G::BR*G::T,&tmIn,full,2*col0,b*G::BR,row);
  };

  if(threadIdx.x == 0) {
    mbarInit(full,1);
    mbarInitFence();
  }
  __syncthreads();
  long long tile=blockIdx.x;
  if(threadIdx.x == 0 && tile < ntiles) issueLoad(tile);

  unsigned phase=0;
  int nfft=0;
  for(; tile < ntiles; tile += gridDim.x) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*G::T;
    mbarWait(full,phase);
    phase ^= 1;
    double2 xin[8];
#pragma unroll
    for(int t=0; t < 8; ++t)
      xin[t]=inS[tileAddr<G::T>(tau+G::TPT*t,lane)];
    __syncthreads(); // every thread holds its inputs: the stage is free
    if(threadIdx.x == 0 && tile+gridDim.x < ntiles) issueLoad(tile+gridDim.x);

    int slot=0;
    for(int isb=0; isb < nsb; ++isb, ++nfft) {
      const SubBlockDev sb=sbs[isb];
      double2 *buf=E0+(nfft & 1)*G::EPAD;
      double2 x[8];
#pragma unroll
      for(int t=0; t < 8; ++t) x[t]=xin[t];
      if(sb.k0 != 0) {
        applyZeta<LG,false>(x,zt+slot*G::ZT,tau);
        ++slot;
      }
      // the bulk store issued two sub-blocks ago read this buffer
      if(threadIdx.x == 0) tmaWaitRead<1>();
      FFT::forward(x,tau,w1,buf,lay);
      __syncthreads(); // last exchange reads done: reuse buf as output tile
#pragma unroll
      for(int e=0; e < 8; ++e) {
        const int l=RegFFT<LG>::rev(8*tau+e);
        buf[tileAddr<G::T>(l,lane)]=x[e];
      }
      fenceAsyncShared();
      __syncthreads();
      if(threadIdx.x == 0) {
#pragma unroll
        for(int b=0; b < G::NBOX; ++b)
          storeSlot(out,isb*G::NBOX+b,buf+b*G::BR*G::T,G::T,2*col0,row);
        tmaCommit();
      }
    }
  }
  if(threadIdx.x == 0) tmaWaitRead<0>();
}

// ---------------------------------------------------------------------------
// backward pass, uniform complex plans with L <= m
// ---------------------------------------------------------------------------
//
// shared memory: [NSTAGE spectrum tiles M*T][E padded][zeta slots][mbarriers]
// One stage per sub-block of a tile (NSTAGE >= nsb is required by the host):
// stage b is refilled with the next tile's sub-block b as soon as every thread
// has taken its 8 points of the current one.
template<int LG, int NSTAGE, int NTHR>
__global__ void __launch_bounds__(NTHR,512/NTHR)
tma_backward_direct(const __grid_constant__ CUtensorMap tmIn,
                    const __grid_constant__ DestSet out, PlanDev P,
                    const SubBlockDev *__restrict__ sbs, int nsb, int layout,
                    double scale, int ntc, long long ntiles, int tabid)
{
  typedef TileGeom<LG,NTHR> G;
  typedef WFFT<LG> FFT;
  extern __shared__ __align__(128) unsigned char smraw[];
  double2 *inS=(double2 *) alignedSmem(smraw);
  double2 *E=inS+NSTAGE*G::TILE;
  double2 *zt=E+G::EPAD;
  int nz=0;
  for(int isb=0; isb < nsb; ++isb) nz += sbs[isb].k0 != 0;
  unsigned long long *full=(unsigned long long *) (zt+nz*G::ZT);

  int lane,tau;
  threadMap<LG,NTHR>(lane,tau);
  PadLaneLayout<G::T> lay;
  lay.lane=lane;
  double2 w1[FFT::NR8 > 0 ? FFT::NR8 : 1];
  FFT::loadW(P.tab[tabid].tw8,tau,w1);
  loadZeta<LG>(P,sbs,nsb,zt);
  const unsigned tileBytes=G::TILE*sizeof(double2);

  auto issueLoad=[&](long long tile, int isb) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*G::T;
    const long long off=layout ? sbs[isb].off_all : sbs[isb].off_call;
    const int r0=(int) (off/P.S);
    mbarExpectTx(full+isb,tileBytes);
#pragma unroll
    for(int b=0; b < G::NBOX; ++b)
      tmaLoad3(inS+isb*G::TILE+b*G::BR*G::T,&tmIn,full+isb,2*col0,r0+b*G::BR,
               row);
  };

  if(threadIdx.x == 0) {
    for(int s=0; s < NSTAGE; ++s) mbarInit(full+s,1);
    mbarInitFence();
  }
  __syncthreads();
  long long tile=blockIdx.x;
  if(threadIdx.x == 0 && tile < ntiles)
    for(int isb=0; isb < nsb; ++isb) issueLoad(tile,isb);

  unsigned phase=0;
  for(; tile < ntiles; tile += gridDim.x) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*G::T;
    double2 racc[8];
#pragma unroll
    for(int t=0; t < 8; ++t) racc[t]=make_double2(0.0,0.0);
    // the previous tile's output store read E
    if(threadIdx.x == 0) tmaWaitRead<0>();
    int slot=0;
    for(int isb=0; isb < nsb; ++isb) {
      const long long k0=sbs[isb].k0;
      mbarWait(full+isb,phase);
      const double2 *st=inS+isb*G::TILE;
      double2 x[8];
#pragma unroll
      for(int e=0; e < 8; ++e) {
        const int l=RegFFT<LG>::rev(8*tau+e);
        x[e]=st[tileAddr<G::T>(l,lane)];
      }
      __syncthreads(); // stage consumed (also orders the wait above for E)
      if(threadIdx.x == 0 && tile+gridDim.x < ntiles)
        issueLoad(tile+gridDim.x,isb);
      FFT::adjoint(x,tau,w1,E,lay);
      if(k0 != 0) {
        applyZeta<LG,true>(x,zt+slot*G::ZT,tau);
        ++slot;
      }
#pragma unroll
      for(int t=0; t < 8; ++t) racc[t]=racc[t]+x[t];
    }
    phase ^= 1;
    __syncthreads(); // exchange reads done: E becomes the output tile
#pragma unroll
    for(int t=0; t < 8; ++t)
      E[tileAddr<G::T>(tau+G::TPT*t,lane)]=wscale(racc[t],scale);
    fenceAsyncShared();
    __syncthreads();
    if(threadIdx.x == 0) {
#pragma unroll
      for(int b=0; b < G::NBOX; ++b)
        storeSlot(out,b,E+b*G::BR*G::T,G::T,2*col0,row);
      tmaCommit();
    }
  }
  if(threadIdx.x == 0) tmaWaitRead<0>();
}

// ---------------------------------------------------------------------------
// real x pass: fftPadReal with p=1, q=2 (the outermost pass of every real
// BASELINE configuration; reference forward1Many/backward1Many,
// convolve.cc:5852-5964,6702-6788)
// ---------------------------------------------------------------------------
//
// Two sub-blocks per tile of T=8 real columns (64-byte rows, 512 of them):
//   r=0   r2c of length M: two adjacent real columns share one complex FFT
//         (z = x_a + i x_b, 4 complex lanes x 64 threads), the spectra are
//         separated with the partner Z[M-l]; e=M/2+1 rows of 8 complex
//         columns (128-byte rows) are stored with the r2c (sign -1) convention;
//   2r=q  packed class: W[s] = zeta^s (x_s + i x_{s+M/2}), complex FFT of
//         length M/2 per column (8 lanes x 32 threads), M/2 rows.
// Shared memory: [input tile 32 KB][X 36 KB][Y 36 KB][zeta][mbarrier].  X and Y
// swap roles every tile so that a buffer is rewritten long after the bulk
// store that read it was issued:
//   r=0 : exchanges + natural-order copy in X, output tile in Y, store(Y)
//   2r=q: exchanges in X, output tile in X, store(X)
template<int LG>
__global__ void __launch_bounds__(256,2)
tma_forward_real(const __grid_constant__ CUtensorMap tmIn,
                 const __grid_constant__ DestSet out, PlanDev P,
                 const SubBlockDev *__restrict__ sbs, int layout, int ntc,
                 long long ntiles, int tab9, int tab8)
{
  typedef TileGeom<LG> G;       // paired block: M, 4 complex lanes
  typedef TileGeom<LG-1> H;     // packed block: M/2, 8 lanes
  typedef WFFT<LG> FFT;
  typedef WFFT<LG-1> FFTH;
  const int M=G::M;
  extern __shared__ __align__(128) unsigned char smraw[];
  double2 *inS=(double2 *) alignedSmem(smraw);   // [M][8 doubles]
  double2 *B0=inS+G::TILE;
  double2 *zt=B0+2*G::EPAD;
  unsigned long long *full=(unsigned long long *) (zt+H::ZT);

  const int cl=threadIdx.x & 3, tau=threadIdx.x >> 2;     // paired mapping
  const int ln=threadIdx.x & 7, tau2=threadIdx.x >> 3;    // packed mapping
  PadLaneLayout<4> lay;
  lay.lane=cl;
  PadLaneLayout<8> layH;
  layH.lane=ln;
  double2 w1[FFT::NR8], w1h[FFTH::NR8];
  FFT::loadW(P.tab[tab9].tw8,tau,w1);
  FFTH::loadW(P.tab[tab8].tw8,tau2,w1h);
  {
    const long long k0=sbs[1].k0;
    for(int j=threadIdx.x; j <= H::TPT; j += blockDim.x)
      zt[j]=zeta(P,modN(P,k0,j));
  }
  const unsigned tileBytes=M*8*sizeof(double);

  auto issueLoad=[&](long long tile) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*8;
    mbarExpectTx(full,tileBytes);
#pragma unroll
    for(int b=0; b < G::NBOX; ++b)
      tmaLoad3(inS+b*G::BR*4,&tmIn,full,col0,b*G::BR,row);
  };

  if(threadIdx.x == 0) {
    mbarInit(full,1);
    mbarInitFence();
  }
  __syncthreads();
  long long tile=blockIdx.x;
  if(threadIdx.x == 0 && tile < ntiles) issueLoad(tile);

  unsigned phase=0;
  int flip=0;
  for(; tile < ntiles; tile += gridDim.x, flip ^= 1) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*8;
    double2 *X=B0+flip*G::EPAD;
    double2 *Y=B0+(flip ^ 1)*G::EPAD;
    mbarWait(full,phase);
    phase ^= 1;

    // ---- r=0: paired r2c ----
    double2 x[8];
#pragma unroll
    for(int t=0; t < 8; ++t) x[t]=inS[(tau+G::TPT*t)*4+cl];
    // X was read by the store issued half a tile ago
    if(threadIdx.x == 0) tmaWaitRead<1>();
    FFT::forward(x,tau,w1,X,lay);
    __syncthreads();
#pragma unroll
    for(int e=0; e < 8; ++e) X[lay.addr(RegFFT<LG>::rev(8*tau+e))]=x[e];
    // Y was read by the store issued at the end of the previous tile
    if(threadIdx.x == 0) tmaWaitRead<0>();
    __syncthreads();
    {
      const int odd=tau & 1;
#pragma unroll
      for(int e=0; e < 8; ++e) {
        const int l=RegFFT<LG>::rev(8*tau+e);
        if(l <= M/2) {
          const double2 z=x[e];
          const double2 zp=X[lay.addr((M-l) & (M-1))];
          // stored with the r2c (sign -1) convention: conj of the + transform
          const double2 xa=make_double2(0.5*(z.x+zp.x),-0.5*(z.y-zp.y));
          const double2 xb=make_double2(0.5*(z.y+zp.y),0.5*(z.x-zp.x));
          // 128-byte rows: even taus write column 2cl first, odd taus column
          // 2cl+1, so that the two rows of a quarter-warp never share a bank
          double2 *o=Y+l*8+2*cl;
          o[odd]=odd ? xb : xa;
          o[odd ^ 1]=odd ? xa : xb;
        }
      }
    }
    fenceAsyncShared();
    __syncthreads();
    if(threadIdx.x == 0) {
      storeSlot(out,0,Y,8,2*col0,row);
      storeSlot(out,1,Y+(M/2)*8,8,2*col0,row);
      tmaCommit();
    }

    // ---- 2r=q: packed class ----
    {
      const double *inD=(const double *) inS;
      double2 z=zt[tau2];
      const double2 step=zt[H::TPT];
#pragma unroll
      for(int t=0; t < 8; ++t) {
        const int s=tau2+H::TPT*t;
        const double2 v=make_double2(inD[s*8+ln],inD[(s+M/2)*8+ln]);
        x[t]=fmul(v,z);
        if(t < 7) z=fmul(z,step);
      }
    }
    __syncthreads(); // both blocks have their inputs: the stage is free
    if(threadIdx.x == 0 && tile+gridDim.x < ntiles) issueLoad(tile+gridDim.x);
    FFTH::forward(x,tau2,w1h,X,layH);
    __syncthreads();
#pragma unroll
    for(int e=0; e < 8; ++e)
      X[RegFFT<LG-1>::rev(8*tau2+e)*8+ln]=x[e];
    fenceAsyncShared();
    __syncthreads();
    if(threadIdx.x == 0) {
      storeSlot(out,2,X,8,2*col0,row);
      tmaCommit();
    }
  }
  if(threadIdx.x == 0) tmaWaitRead<0>();
}

// Backward real x pass.  Shared memory:
//   [A: r=0 spectrum (M/2+1) x 8 complex][Bs: packed spectrum M/2 x 8 complex]
//   [E 36 KB][zeta][2 mbarriers]
// The packed class runs first and leaves its contribution 2 Re / 2 Im in E as
// a real tile; the paired threads take it into their accumulators, run the
// c2r transform of the r=0 block and write the finished tile for the store.
template<int LG>
__global__ void __launch_bounds__(256,2)
tma_backward_real(const __grid_constant__ CUtensorMap tmIn,
                  const __grid_constant__ CUtensorMap tmIn1,
                  const __grid_constant__ CUtensorMap tmOut, PlanDev P,
                  const SubBlockDev *__restrict__ sbs, int layout,
                  double scale, int ntc, long long ntiles, int tab9, int tab8)
{
  typedef TileGeom<LG> G;
  typedef TileGeom<LG-1> H;
  typedef WFFT<LG> FFT;
  typedef WFFT<LG-1> FFTH;
  const int M=G::M;
  const int ASZ=(M/2+8)*8;       // double2 words reserved for stage A
  extern __shared__ __align__(128) unsigned char smraw[];
  double2 *A=(double2 *) alignedSmem(smraw);
  double2 *Bs=A+ASZ;
  double2 *E=Bs+(M/2)*8;
  double2 *zt=E+G::EPAD;
  unsigned long long *full=(unsigned long long *) (zt+H::ZT);

  const int cl=threadIdx.x & 3, tau=threadIdx.x >> 2;
  const int ln=threadIdx.x & 7, tau2=threadIdx.x >> 3;
  PadLaneLayout<4> lay;
  lay.lane=cl;
  PadLaneLayout<8> layH;
  layH.lane=ln;
  double2 w1[FFT::NR8], w1h[FFTH::NR8];
  FFT::loadW(P.tab[tab9].tw8,tau,w1);
  FFTH::loadW(P.tab[tab8].tw8,tau2,w1h);
  {
    const long long k0=sbs[1].k0;
    for(int j=threadIdx.x; j <= H::TPT; j += blockDim.x)
      zt[j]=zeta(P,modN(P,k0,j));
  }
  const int r0a=(int) ((layout ? sbs[0].off_all : sbs[0].off_call)/P.S);
  const int r0b=(int) ((layout ? sbs[1].off_all : sbs[1].off_call)/P.S);
  const unsigned rowBytes=8*sizeof(double2);

  auto issueA=[&](long long tile) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*8;
    mbarExpectTx(full,(M/2+1)*rowBytes);
    tmaLoad3(A,&tmIn,full,2*col0,r0a,row);
    tmaLoad3(A+(M/2)*8,&tmIn1,full,2*col0,r0a+M/2,row);
  };
  auto issueB=[&](long long tile) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*8;
    mbarExpectTx(full+1,(M/2)*rowBytes);
    tmaLoad3(Bs,&tmIn,full+1,2*col0,r0b,row);
  };

  if(threadIdx.x == 0) {
    mbarInit(full,1);
    mbarInit(full+1,1);
    mbarInitFence();
  }
  __syncthreads();
  long long tile=blockIdx.x;
  if(threadIdx.x == 0 && tile < ntiles) {
    issueB(tile);
    issueA(tile);
  }

  unsigned phase=0;
  for(; tile < ntiles; tile += gridDim.x) {
    const int row=(int) (tile/ntc);
    const int col0=(int) (tile % ntc)*8;
    double2 x[8];

    // ---- packed class ----
    mbarWait(full+1,phase);
#pragma unroll
    for(int e=0; e < 8; ++e)
      x[e]=Bs[RegFFT<LG-1>::rev(8*tau2+e)*8+ln];
    __syncthreads();
    if(threadIdx.x == 0) {
      if(tile+gridDim.x < ntiles) issueB(tile+gridDim.x);
      tmaWaitRead<0>(); // the previous tile's store read E
    }
    FFTH::adjoint(x,tau2,w1h,E,layH);
    {
      double2 z=zt[tau2];
      const double2 step=zt[H::TPT];
#pragma unroll
      for(int t=0; t < 8; ++t) {
        x[t]=fmulc(x[t],z);
        if(t < 7) z=fmul(z,step);
      }
    }
    __syncthreads(); // exchange reads done
    {
      double *Ed=(double *) E;
#pragma unroll
      for(int t=0; t < 8; ++t) {
        const int s=tau2+H::TPT*t;
        Ed[s*8+ln]=2.0*x[t].x;
        Ed[(s+M/2)*8+ln]=2.0*x[t].y;
      }
    }
    __syncthreads();

    // ---- r=0: paired c2r ----
    double2 racc[8];
#pragma unroll
    for(int t=0; t < 8; ++t) racc[t]=E[(tau+G::TPT*t)*4+cl];
    mbarWait(full,phase);
    {
      const int odd=tau & 1;
#pragma unroll
      for(int e=0; e < 8; ++e) {
        const int l=RegFFT<LG>::rev(8*tau+e);
        // G[l] of the + transform: conj(stored[l]) for l <= M/2, else
        // stored[M-l]
        const bool lower=l <= M/2;
        const double2 *q=A+(lower ? l : M-l)*8+2*cl;
        const double2 q0=q[odd], q1=q[odd ^ 1];
        double2 ga=odd ? q1 : q0, gb=odd ? q0 : q1;
        if(lower) {ga.y=-ga.y; gb.y=-gb.y;}
        x[e]=make_double2(ga.x-gb.y,ga.y+gb.x);
      }
    }
    __syncthreads(); // stage A consumed; packed contribution taken from E
    if(threadIdx.x == 0 && tile+gridDim.x < ntiles) issueA(tile+gridDim.x);
    phase ^= 1;
    FFT::adjoint(x,tau,w1,E,lay);
#pragma unroll
    for(int t=0; t < 8; ++t)
      racc[t]=make_double2(racc[t].x+x[t].x,racc[t].y+x[t].y);
    __syncthreads();
#pragma unroll
    for(int t=0; t < 8; ++t) E[(tau+G::TPT*t)*4+cl]=wscale(racc[t],scale);
    fenceAsyncShared();
    __syncthreads();
    if(threadIdx.x == 0) {
#pragma unroll
      for(int b=0; b < G::NBOX; ++b)
        tmaStore3(&tmOut,E+b*G::BR*4,col0,b*G::BR,row);
      tmaCommit();
    }
  }
  if(threadIdx.x == 0) tmaWaitRead<0>();
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType,
                                  cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled()
{
  static EncodeTiledFn fn=NULL;
  static bool tried=false;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if(!tried) {
    tried=true;
    void *p=NULL;
    cudaDriverEntryPointQueryResult q;
    if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&p,cudaEnableDefault,
                               &q) == cudaSuccess &&
       q == cudaDriverEntryPointSuccess)
      fn=(EncodeTiledFn) p;
    cudaGetLastError();
  }
  return fn;
}

// 3-D tensor of doubles: dim0 contiguous (n0 doubles), dim1 n1 rows s1 bytes
// apart, dim2 n2 planes s2 bytes apart; box b0 x b1 x 1.  Out-of-bounds box
// elements read as zero and are not written.
// FFTWPP_TMA_L2PROMO (experiment builds): 0 none, 1 64 B, 2 128 B, 3 256 B
int l2Promotion()
{
#ifdef FFTWPP_EXPERIMENT_SWITCHES
  static int v=-1;
  if(v < 0) {
    const char *s=getenv("FFTWPP_TMA_L2PROMO");
    v=s ? atoi(s) : FFTWPP_TMA_L2PROMO_DEFAULT;
    if(v < 0 || v > 3) v=0;
  }
  return v;
#else
  return FFTWPP_TMA_L2PROMO_DEFAULT;
#endif
}

bool makeMap(CUtensorMap *map, const void *base, uint64_t n0, uint64_t n1,
             uint64_t s1, uint64_t n2, uint64_t s2, uint32_t b0, uint32_t b1)
{
  EncodeTiledFn enc=encodeTiled();
  if(!enc) return false;
  if(((uintptr_t) base & 15) || (s1 & 15) || (s2 & 15) || n0 == 0 || n1 == 0 ||
     n2 == 0)
    return false;
  if(n0 >= (1ull << 32) || n1 >= (1ull << 32) || n2 >= (1ull << 32) ||
     s1 >= (1ull << 40) || s2 >= (1ull << 40))
    return false;
  cuuint64_t dim[3]={n0,n1,n2};
  cuuint64_t stride[2]={s1,s2};
  cuuint32_t box[3]={b0,b1,1};
  cuuint32_t es[3]={1,1,1};
  CUresult r=enc(map,CU_TENSOR_MAP_DATA_TYPE_FLOAT64,3,(void *) base,dim,
                 stride,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE,
                 (CUtensorMapL2promotion) l2Promotion(),
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// FFTWPP_NO_TMA=1 keeps every strided pass on the register-gather kernels of
// fast_kernels.cu (testing and A/B timing); FFTWPP_NO_TMA_REAL=1 only the real
// x pass, and only in builds with -DFFTWPP_EXPERIMENT_SWITCHES.
bool tmaRealDisabled()
{
#ifdef FFTWPP_EXPERIMENT_SWITCHES
  static int off=-1;
  if(off < 0) {
    const char *s=getenv("FFTWPP_NO_TMA_REAL");
    off=(s && *s && *s != '0') ? 1 : 0;
  }
  return off == 1;
#else
  return false;
#endif
}

// 8-lane (128-byte row) tiles for remote destinations: measured 2.31 vs 2.51
// ms per convolution at N=8.  FFTWPP_TMA_WIDE_REMOTE=0 (experiment builds)
// keeps the 4-lane tiles for A/B timing.
bool wideRemoteTiles()
{
#ifdef FFTWPP_EXPERIMENT_SWITCHES
  static int on=-1;
  if(on < 0) {
    const char *s=getenv("FFTWPP_TMA_WIDE_REMOTE");
    on=(s && *s == '0') ? 0 : 1;
  }
  return on == 1;
#else
  return true;
#endif
}

bool tmaDisabled()
{
  static int off=-1;
  if(off < 0) {
    const char *s=getenv("FFTWPP_NO_TMA");
    off=(s && *s && *s != '0') ? 1 : 0;
  }
  return off == 1;
}

template<class K>
int allowSmemTma(K kernel, size_t bytes, int budget=113*1024)
{
  static std::mutex mu;
  static std::vector<std::pair<const void *,int> > done;
  int dev=0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for(size_t i=0; i < done.size(); ++i)
    if(done[i].first == (const void *) kernel && done[i].second == dev)
      return 0;
  // the budget of two CTAs per SM, whatever this particular launch needs
  // (the number of residue-twiddle slots varies between launches)
  (void) bytes;
  cudaError_t e=cudaFuncSetAttribute(kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     budget);
  if(e != cudaSuccess) return cuda_fail(e,"cudaFuncSetAttribute");
  done.push_back(std::make_pair((const void *) kernel,dev));
  return 0;
}

// Common eligibility of the direct (uniform complex, one term) passes.
// On success fills rows (all-layout / call-layout row extent) and nz.
bool directEligible(Plan *pl, uint64_t sb0, uint64_t nsb, int layout, int lg,
                    uint64_t& rowsMax, int& nz)
{
  FastInfo *fi=pl->fast;
  const PlanDev& d=pl->dev;
  if(!fi || !fi->uniform || fi->nterm != 1) return false;
  if(d.kind != FFTWPP_KIND_COMPLEX || d.jmin != 0 || d.oen) return false;
  if(lg != 9) return false; // instantiated tile shapes (T=4 lanes x 64 threads)
  const int M=1 << lg;
  if(d.Lin > M || d.C < 4) return false;
  rowsMax=0;
  nz=0;
  for(uint64_t i=sb0; i < sb0+nsb; ++i) {
    const SubBlockDev& sb=pl->hsub[i];
    if((int) sb.mlen != M || (int) sb.nout != M || sb.flags) return false;
    const long long off=layout ? sb.off_all : sb.off_call;
    if(off % d.S) return false;
    rowsMax=std::max<uint64_t>(rowsMax,(uint64_t) (off/d.S)+sb.nout);
    nz += sb.k0 != 0;
  }
  int tabid=-1;
  for(int k=0; k < 2; ++k)
    if(d.tab[k].n == M) tabid=k;
  return tabid >= 0;
}

int tableId(const PlanDev& d, int M)
{
  for(int k=0; k < 2; ++k)
    if(d.tab[k].n == M) return k;
  return -1;
}

// fftPadReal, p=1, q=2, m=2^lg: the r2c block followed by the packed class
bool realEligible(Plan *pl, uint64_t sb0, uint64_t nsb, int layout, int lg,
                  uint64_t& rowsMax)
{
  FastInfo *fi=pl->fast;
  const PlanDev& d=pl->dev;
  if(!fi || d.kind != FFTWPP_KIND_REAL || lg != 9 || fi->log2m != lg)
    return false;
  const int M=1 << lg;
  if(sb0 != 0 || nsb != 2 || pl->hsub.size() != 2) return false;
  if(d.N != 2*M || d.jmin != 0 || d.Lin > M || d.oen) return false;
  if(d.C < 8 || (d.C & 1) || (d.S & 1)) return false;
  const SubBlockDev& a=pl->hsub[0];
  const SubBlockDev& b=pl->hsub[1];
  if((int) a.mlen != M || (int) a.nout != M/2+1 ||
     a.flags != FFTWPP_SB_CONJ_OUT || a.k0 != 0)
    return false;
  if((int) b.mlen != M/2 || (int) b.nout != M/2 || b.flags != 0 || b.k0 != 1)
    return false;
  rowsMax=0;
  for(int i=0; i < 2; ++i) {
    const SubBlockDev& sb=pl->hsub[i];
    const long long off=layout ? sb.off_all : sb.off_call;
    if(off % d.S) return false;
    rowsMax=std::max<uint64_t>(rowsMax,(uint64_t) (off/d.S)+sb.nout);
  }
  return tableId(d,M) >= 0 && tableId(d,M/2) >= 0;
}

// One store slot of a kernel: `rows` output rows starting at absolute row
// `row0` (all-layout row of a forward pass, input index j of a backward pass).
struct Slot {
  uint64_t row0, rows;
};

// Cut the slots into per-owner segments and encode their tensor maps.
// dests == NULL: the dense local layout (rows S words apart, planes
// planeStride words apart).  rowBytes: bytes of one staged tile row (segment
// sources must stay 128-byte aligned in shared memory).
bool makeDests(DestSet *D, const fftwpp_gpu_dest *dests, int ndest,
               void *local, uint64_t C, uint64_t S, uint64_t rowsMax,
               uint64_t nplanes, uint64_t planeStride, uint64_t plane0,
               uint32_t boxCols, size_t rowBytes, const Slot *slots,
               int nslots)
{
  const uint64_t w=sizeof(double2);
  fftwpp_gpu_dest one;
  if(!dests) {
    one.base=local;
    one.row0=0;
    one.rows=rowsMax;
    one.row_stride=S;
    one.plane_stride=planeStride;
    dests=&one;
    ndest=1;
    plane0=0;
  }
  if(ndest < 1 || nslots > MAXSLOT) return false;
  D->plane0=(int) plane0;
  const uint64_t planes=plane0+nplanes;
  int nseg=0;
  for(int sl=0; sl < nslots; ++sl) {
    D->first[sl]=nseg;
    const uint64_t lo=slots[sl].row0, hi=lo+slots[sl].rows;
    uint64_t covered=0;
    for(int p=0; p < ndest; ++p) {
      const fftwpp_gpu_dest& t=dests[p];
      const uint64_t a=std::max<uint64_t>(lo,t.row0);
      const uint64_t b=std::min<uint64_t>(hi,t.row0+t.rows);
      if(a >= b) continue;
      if(nseg >= MAXSEG) return false;
      if(((a-lo)*rowBytes) & 127) return false;
      const uint64_t s2=(planes > 1 ? t.plane_stride : t.row_stride*t.rows)*w;
      if(!makeMap(&D->map[nseg],t.base,2*C,t.rows,t.row_stride*w,planes,s2,
                  boxCols,(uint32_t) (b-a)))
        return false;
      D->srcRow[nseg]=(int) (a-lo);
      D->dstRow[nseg]=(int) (a-t.row0);
      covered += b-a;
      ++nseg;
    }
    if(covered != hi-lo) return false; // every output row needs an owner
  }
  for(int sl=nslots; sl <= MAXSLOT; ++sl) D->first[sl]=nseg;
  return true;
}

} // namespace

int tma_try_forward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                    const void *f, void *F, uint64_t nrows, uint64_t frs,
                    uint64_t Frs, cudaStream_t st,
                    const fftwpp_gpu_dest *dests, int ndest, uint64_t plane0)
{
  if(tmaDisabled() || !pl->fast) return 0;
  const int lg=pl->fast->log2m;
  uint64_t rowsMax;
  int nz;
  typedef TileGeom<9> G;
  const PlanDev& d=pl->dev;
  const uint64_t w=sizeof(double2);
  if(dests && !layout) return 0; // destinations own ranges of all-layout rows
  if(realEligible(pl,sb0,nsb,layout,lg,rowsMax) && !tmaRealDisabled()) {
    typedef TileGeom<8> H;
    if(nrows == 0) return 1;
    if(nrows > 1 && ((frs & 1) || (!dests && Frs < rowsMax*(uint64_t) d.S)))
      return 0;
    CUtensorMap tmIn;
    DestSet out;
    if(!makeMap(&tmIn,f,(uint64_t) d.C,(uint64_t) d.Lin,(uint64_t) d.S*8,nrows,
                (nrows > 1 ? frs : (uint64_t) d.S*d.Lin)*8,8,G::BR))
      return 0;
    {
      const uint64_t ra=(uint64_t) (pl->hsub[0].off_all/d.S);
      const uint64_t rb=(uint64_t) (pl->hsub[1].off_all/d.S);
      const uint64_t ca=(uint64_t) (pl->hsub[0].off_call/d.S);
      const uint64_t cb=(uint64_t) (pl->hsub[1].off_call/d.S);
      const Slot slots[3]={{layout ? ra : ca,(uint64_t) G::M/2},
                           {(layout ? ra : ca)+G::M/2,1},
                           {layout ? rb : cb,(uint64_t) G::M/2}};
      if(!makeDests(&out,dests,ndest,F,d.C,d.S,rowsMax,nrows,
                    nrows > 1 ? Frs : (uint64_t) d.S*rowsMax,plane0,16,128,
                    slots,3))
        return 0;
    }
    const size_t smem=(size_t) (G::TILE+2*G::EPAD+H::ZT)*w+16+1024;
    if(smem > 113*1024) return 0;
    const int ntc=(d.C+7)/8;
    const long long ntiles=(long long) nrows*ntc;
    const unsigned grid=(unsigned) std::min<long long>(ntiles,
                                                       (long long) sm_count()*2);
    int rc=allowSmemTma(tma_forward_real<9>,smem);
    if(rc) return rc;
    prof_begin(4*pl->tag+0,st);
    tma_forward_real<9><<<grid,G::NT,smem,st>>>
      (tmIn,out,pl->dev,pl->dsub,layout,ntc,ntiles,tableId(d,G::M),
       tableId(d,G::M/2));
    rc=check_launch("tma_forward_real",st);
    return rc ? rc : 1;
  }
  if(!directEligible(pl,sb0,nsb,layout,lg,rowsMax,nz)) return 0;
  if(nrows == 0) return 1;
  if(nrows > 1 && !dests && Frs < rowsMax*(uint64_t) d.S) return 0;
  CUtensorMap tmIn;
  DestSet out;
  // input: C columns, Lin rows S words apart, nrows planes frs words apart
  if(!makeMap(&tmIn,f,2*(uint64_t) d.C,(uint64_t) d.Lin,(uint64_t) d.S*w,nrows,
              (nrows > 1 ? frs : (uint64_t) d.S*d.Lin)*w,2*G::T,G::BR))
    return 0;
  {
    Slot slots[MAXSLOT];
    if(nsb*G::NBOX > (uint64_t) MAXSLOT) return 0;
    int ns=0;
    for(uint64_t i=sb0; i < sb0+nsb; ++i) {
      const SubBlockDev& sb=pl->hsub[i];
      const uint64_t r0=(uint64_t) ((layout ? sb.off_all : sb.off_call)/d.S);
      for(int b=0; b < G::NBOX; ++b) {
        slots[ns].row0=r0+b*G::BR;
        slots[ns].rows=G::BR;
        ++ns;
      }
    }
    if(!makeDests(&out,dests,ndest,F,d.C,d.S,rowsMax,nrows,
                  nrows > 1 ? Frs : (uint64_t) d.S*rowsMax,plane0,2*G::T,
                  G::T*sizeof(double2),slots,ns))
      return 0;
  }
  const size_t smem=(size_t) (G::TILE+2*G::EPAD+nz*G::ZT)*w+16+1024;
  if(smem > 113*1024) return 0;
  const int ntc=(d.C+G::T-1)/G::T;
  const long long ntiles=(long long) nrows*ntc;
  const unsigned grid=(unsigned) std::min<long long>(ntiles,
                                                     (long long) sm_count()*2);
  int rc=allowSmemTma(tma_forward_direct<9>,smem);
  if(rc) return rc;
  prof_begin(4*pl->tag+0,st);
  tma_forward_direct<9><<<grid,G::NT,smem,st>>>
    (tmIn,out,pl->dev,pl->dsub+sb0,(int) nsb,layout,ntc,ntiles,
     tableId(d,G::M));
  rc=check_launch("tma_forward_direct",st);
  return rc ? rc : 1;
}

int tma_try_backward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                     const void *F, void *f, int accumulate, double scale,
                     uint64_t nrows, uint64_t Frs, uint64_t frs,
                     cudaStream_t st, const fftwpp_gpu_dest *dests, int ndest,
                     uint64_t plane0)
{
  if(tmaDisabled() || !pl->fast || accumulate) return 0;
  const int lg=pl->fast->log2m;
  uint64_t rowsMax;
  int nz;
  typedef TileGeom<9> G;
  const PlanDev& d=pl->dev;
  const uint64_t w=sizeof(double2);
  if(!dests && realEligible(pl,sb0,nsb,layout,lg,rowsMax) &&
     !tmaRealDisabled()) {
    typedef TileGeom<8> H;
    if(nrows == 0) return 1;
    if(nrows > 1 && ((frs & 1) || Frs < rowsMax*(uint64_t) d.S)) return 0;
    CUtensorMap tmIn,tmIn1,tmOut;
    const uint64_t s2=(nrows > 1 ? Frs : (uint64_t) d.S*rowsMax)*w;
    if(!makeMap(&tmIn,F,2*(uint64_t) d.C,rowsMax,(uint64_t) d.S*w,nrows,s2,16,
                G::BR) ||
       !makeMap(&tmIn1,F,2*(uint64_t) d.C,rowsMax,(uint64_t) d.S*w,nrows,s2,
                16,1))
      return 0;
    if(!makeMap(&tmOut,f,(uint64_t) d.C,(uint64_t) d.Lin,(uint64_t) d.S*8,
                nrows,(nrows > 1 ? frs : (uint64_t) d.S*d.Lin)*8,8,G::BR))
      return 0;
    const size_t smem=(size_t) ((G::M/2+8)*8+(G::M/2)*8+G::EPAD+H::ZT)*w+32+
      1024;
    if(smem > 113*1024) return 0;
    const int ntc=(d.C+7)/8;
    const long long ntiles=(long long) nrows*ntc;
    const unsigned grid=(unsigned) std::min<long long>(ntiles,
                                                       (long long) sm_count()*2);
    int rc=allowSmemTma(tma_backward_real<9>,smem);
    if(rc) return rc;
    prof_begin(4*pl->tag+1,st);
    tma_backward_real<9><<<grid,G::NT,smem,st>>>
      (tmIn,tmIn1,tmOut,pl->dev,pl->dsub,layout,scale,ntc,ntiles,
       tableId(d,G::M),tableId(d,G::M/2));
    rc=check_launch("tma_backward_real",st);
    return rc ? rc : 1;
  }
  if(!directEligible(pl,sb0,nsb,layout,lg,rowsMax,nz)) return 0;
  if(nsb > 2) return 0; // one staged tile per sub-block, two CTAs per SM
  if(nrows == 0) return 1;
  if(nrows > 1 && (Frs < rowsMax*(uint64_t) d.S)) return 0;
  // Remote destinations (fused exchange): tiles of 8 lanes, so that the rows
  // crossing NVLink are 128 bytes long (64-byte rows reach 392 GB/s, 128-byte
  // rows 633 GB/s at N=8); one CTA of 512 threads per SM -- the pass is then
  // bound by the link, not by the SM.
  if(dests && ndest > 1 && d.C >= 8 && wideRemoteTiles()) {
    typedef TileGeom<9,512> W;
    CUtensorMap tmInW;
    DestSet outW;
    if(makeMap(&tmInW,F,2*(uint64_t) d.C,rowsMax,(uint64_t) d.S*w,nrows,
               (nrows > 1 ? Frs : (uint64_t) d.S*rowsMax)*w,2*W::T,W::BR)) {
      Slot slots[W::NBOX];
      int ns=0;
      for(int b=0; b < W::NBOX; ++b) {
        const uint64_t lo=(uint64_t) b*W::BR;
        const uint64_t hi=std::min<uint64_t>(lo+W::BR,(uint64_t) d.Lin);
        slots[ns].row0=lo;
        slots[ns].rows=hi > lo ? hi-lo : 0;
        ++ns;
      }
      const size_t smemW=(size_t) (2*W::TILE+W::EPAD+nz*W::ZT)*w+32+1024;
      if(smemW <= 227*1024 &&
         makeDests(&outW,dests,ndest,f,d.C,d.S,(uint64_t) d.Lin,nrows,
                   nrows > 1 ? frs : (uint64_t) d.S*d.Lin,plane0,2*W::T,
                   W::T*sizeof(double2),slots,ns)) {
        const int ntcw=(d.C+W::T-1)/W::T;
        const long long ntilesw=(long long) nrows*ntcw;
        const unsigned gridw=(unsigned) std::min<long long>(ntilesw,
                                                            (long long) sm_count());
        int rcw=allowSmemTma(tma_backward_direct<9,2,512>,smemW,227*1024);
        if(rcw) return rcw;
        prof_begin(4*pl->tag+1,st);
        tma_backward_direct<9,2,512><<<gridw,W::NT,smemW,st>>>
          (tmInW,outW,pl->dev,pl->dsub+sb0,(int) nsb,layout,scale,ntcw,ntilesw,
           tableId(d,W::M));
        rcw=check_launch("tma_backward_direct (wide)",st);
        return rcw ? rcw : 1;
      }
    }
  }
  CUtensorMap tmIn;
  DestSet out;
  if(!makeMap(&tmIn,F,2*(uint64_t) d.C,rowsMax,(uint64_t) d.S*w,nrows,
              (nrows > 1 ? Frs : (uint64_t) d.S*rowsMax)*w,2*G::T,G::BR))
    return 0;
  // output rows are the input indices j of the transformed dimension
  {
    // the tile holds M rows; rows j >= Lin do not exist in the output
    Slot slots[G::NBOX];
    int ns=0;
    for(int b=0; b < G::NBOX; ++b) {
      const uint64_t lo=(uint64_t) b*G::BR;
      const uint64_t hi=std::min<uint64_t>(lo+G::BR,(uint64_t) d.Lin);
      slots[ns].row0=lo;
      slots[ns].rows=hi > lo ? hi-lo : 0;
      ++ns;
    }
    if(!makeDests(&out,dests,ndest,f,d.C,d.S,(uint64_t) d.Lin,nrows,
                  nrows > 1 ? frs : (uint64_t) d.S*d.Lin,plane0,2*G::T,
                  G::T*sizeof(double2),slots,ns))
      return 0;
  }
  const size_t smem=(size_t) (2*G::TILE+G::EPAD+nz*G::ZT)*w+32+1024;
  if(smem > 113*1024) return 0;
  const int ntc=(d.C+G::T-1)/G::T;
  const long long ntiles=(long long) nrows*ntc;
  const unsigned grid=(unsigned) std::min<long long>(ntiles,
                                                     (long long) sm_count()*2);
  int rc=allowSmemTma(tma_backward_direct<9,2,256>,smem);
  if(rc) return rc;
  prof_begin(4*pl->tag+1,st);
  tma_backward_direct<9,2,256><<<grid,G::NT,smem,st>>>
    (tmIn,out,pl->dev,pl->dsub+sb0,(int) nsb,layout,scale,ntc,ntiles,
     tableId(d,G::M));
  rc=check_launch("tma_backward_direct",st);
  return rc ? rc : 1;
}

} // namespace fftwpp_gpu
