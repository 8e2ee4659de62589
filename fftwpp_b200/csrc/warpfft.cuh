// warpfft.cuh -- one-warp transforms of length 512: 16 points per thread,
// 512 = 16 x 16 x 2: two in-register 16-point DFTs with ONE shared-memory
// exchange between them and the final radix-2 step across lane pairs by warp
// shuffles; no CTA-wide or named barriers.  The result is in a scrambled order
// that the fused convolutions never undo (adjoint() is the exact adjoint).
#ifndef FFTWPP_WARPFFT_CUH
#define FFTWPP_WARPFFT_CUH

#include "regfft.cuh"

namespace fftwpp_gpu {
namespace {

// omega_32^k, k < 16 (after unrolling k is a compile-time constant)
__device__ __forceinline__ double2 w32(int k)
{
  const double c1=0.98078528040323044913, s1=0.19509032201612826785;
  const double c2=0.92387953251128675613, s2=0.38268343236508977173;
  const double c3=0.83146961230254523708, s3=0.55557023301960222474;
  const double h=0.70710678118654752440;
  switch(k) {
    case 0: return make_double2(1.0,0.0);
    case 1: return make_double2(c1,s1);
    case 2: return make_double2(c2,s2);
    case 3: return make_double2(c3,s3);
    case 4: return make_double2(h,h);
    case 5: return make_double2(s3,c3);
    case 6: return make_double2(s2,c2);
    case 7: return make_double2(s1,c1);
    case 8: return make_double2(0.0,1.0);
    case 9: return make_double2(-s1,c1);
    case 10: return make_double2(-s2,c2);
    case 11: return make_double2(-s3,c3);
    case 12: return make_double2(-h,h);
    case 13: return make_double2(-c3,s3);
    case 14: return make_double2(-c2,s2);
    default: return make_double2(-c1,s1);
  }
}

// a*omega_16^{SIGN e} for the exponents of a 4x4 decomposition
template<int SIGN>
__device__ __forceinline__ double2 mulw16(double2 a, int e)
{
  const double h=0.70710678118654752440;
  if(e == 0) return a;
  if(e == 4) return rot<SIGN>(a);
  if(e == 2) { // h (1 + SIGN i)
    double2 r=rot<SIGN>(a);
    return make_double2(h*(a.x+r.x),h*(a.y+r.y));
  }
  if(e == 6) { // h (-1 + SIGN i)
    double2 r=rot<SIGN>(a);
    return make_double2(h*(r.x-a.x),h*(r.y-a.y));
  }
  double2 w=w32((2*e) & 15);
  if(e == 9) w=make_double2(-w32(2).x,-w32(2).y);
  return SIGN > 0 ? fmul(a,w) : fmulc(a,w);
}

// 16-point DFT in registers, 4 x 4: input a[t], t=4 t1+t0; output register
// r=4 k0+k1 holds index k=k0+4 k1 (see K16).
template<int SIGN>
__device__ __forceinline__ void dft16_fwd(double2 (&a)[16])
{
#pragma unroll
  for(int t0=0; t0 < 4; ++t0) bfly4<SIGN>(a[t0],a[4+t0],a[8+t0],a[12+t0]);
#pragma unroll
  for(int k0=1; k0 < 4; ++k0)
#pragma unroll
    for(int t0=1; t0 < 4; ++t0)
      a[4*k0+t0]=mulw16<SIGN>(a[4*k0+t0],k0*t0);
#pragma unroll
  for(int k0=0; k0 < 4; ++k0)
    bfly4<SIGN>(a[4*k0],a[4*k0+1],a[4*k0+2],a[4*k0+3]);
}

// the same three steps in reverse order: dft16_rev<-S> is the adjoint of
// dft16_fwd<S> (input register 4 k0+k1 holds index k0+4 k1, output a[t])
template<int SIGN>
__device__ __forceinline__ void dft16_rev(double2 (&a)[16])
{
#pragma unroll
  for(int k0=0; k0 < 4; ++k0)
    bfly4<SIGN>(a[4*k0],a[4*k0+1],a[4*k0+2],a[4*k0+3]);
#pragma unroll
  for(int k0=1; k0 < 4; ++k0)
#pragma unroll
    for(int t0=1; t0 < 4; ++t0)
      a[4*k0+t0]=mulw16<SIGN>(a[4*k0+t0],k0*t0);
#pragma unroll
  for(int t0=0; t0 < 4; ++t0) bfly4<SIGN>(a[t0],a[4+t0],a[8+t0],a[12+t0]);
}

__device__ __forceinline__ double2 shflx1(double2 v)
{
  return make_double2(__shfl_xor_sync(0xffffffffu,v.x,1),
                      __shfl_xor_sync(0xffffffffu,v.y,1));
}

struct WarpFFT512 {
  static const int ROW=34;        // padded row of the 16 x 32 exchange buffer
  static const int BUF=16*ROW;
  static __host__ __device__ constexpr int K16(int r) {return (r >> 2)+4*(r & 3);}

  // in: x[t]=W[lane+32 t]; twa[k*32+l]=omega_512^{k l}; buf: BUF words owned
  // by this warp.  out: scrambled spectrum (16 values per lane)
  static __device__ __forceinline__ void forward(double2 (&x)[16], int lane,
                                                 const double2 *twa,
                                                 double2 *buf) {
    dft16_fwd<1>(x);
#pragma unroll
    for(int r=1; r < 16; ++r) x[r]=fmul(x[r],twa[K16(r)*32+lane]);
    __syncwarp();
#pragma unroll
    for(int r=0; r < 16; ++r) buf[K16(r)*ROW+lane]=x[r];
    __syncwarp();
    const int g=lane >> 1;
    const int c=lane & 1;
#pragma unroll
    for(int t=0; t < 16; ++t) x[t]=buf[g*ROW+c+2*t];
    dft16_fwd<1>(x);
#pragma unroll
    for(int r=1; r < 16; ++r) {
      double2 t=fmul(x[r],w32(K16(r)));
      if(c) x[r]=t;
    }
#pragma unroll
    for(int j=0; j < 8; ++j) {
      double2 recv=shflx1(c ? x[2*j] : x[2*j+1]);
      double2 a=c ? recv : x[2*j];
      double2 b=c ? x[2*j+1] : recv;
      x[2*j]=a+b;
      x[2*j+1]=a-b;
    }
  }

  // exact adjoint of forward(): out x[t]=w[lane+32 t]
  static __device__ __forceinline__ void adjoint(double2 (&x)[16], int lane,
                                                 const double2 *twa,
                                                 double2 *buf) {
    const int g=lane >> 1;
    const int c=lane & 1;
#pragma unroll
    for(int j=0; j < 8; ++j) {
      double2 u=x[2*j]+x[2*j+1];
      double2 v=x[2*j]-x[2*j+1];
      double2 recv=shflx1(c ? u : v);
      x[2*j]=c ? recv : u;
      x[2*j+1]=c ? v : recv;
    }
#pragma unroll
    for(int r=1; r < 16; ++r) {
      double2 t=fmulc(x[r],w32(K16(r)));
      if(c) x[r]=t;
    }
    dft16_rev<-1>(x);
    __syncwarp();
#pragma unroll
    for(int t=0; t < 16; ++t) buf[g*ROW+c+2*t]=x[t];
    __syncwarp();
#pragma unroll
    for(int r=0; r < 16; ++r) x[r]=buf[K16(r)*ROW+lane];
#pragma unroll
    for(int r=1; r < 16; ++r) x[r]=fmulc(x[r],twa[K16(r)*32+lane]);
    dft16_rev<-1>(x);
  }
};

} // namespace
} // namespace fftwpp_gpu

#endif
