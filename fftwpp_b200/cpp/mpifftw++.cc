// mpifftw++.cc -- see mpifftw++.h.
#include "mpifftw++.h"

#include <iostream>

using namespace utils;

namespace fftwpp {

namespace {

// explicit (q=1) plan: the plain DFT of length N over C interleaved columns
// (NULL for the identity transform N == 1)
fftBase *plainPlan(size_t N, Application& app, size_t C)
{
  return N > 1 ? new fftPad(N,N,app,C,C,N,1,true) : NULL;
}

// sign +1: the plan's forward pass; sign -1: its adjoint
// `words`: Complex words the pass covers (for the identity transform)
void transform(fftBase *fft, int sgn, const void *in, void *out, size_t nrows,
               size_t rowstride, size_t words)
{
  if(nrows == 0 || words == 0) return;
  void *st=gpu::stream();
  if(!fft) {
    if(in != out)
      gpu::check(fftwpp_gpu_memcpy_d2d(out,in,words*sizeof(Complex),st),
                 "copy");
    return;
  }
  if(sgn > 0)
    gpu::check(fftwpp_gpu_forward(fft->plan(),0,1,1,in,out,nrows,rowstride,
                                  rowstride,st),"forward (fft)");
  else
    gpu::check(fftwpp_gpu_backward(fft->plan(),0,1,1,in,out,0,1.0,nrows,
                                   rowstride,rowstride,st),"backward (fft)");
}

void scaleDoubles(void *f, double s, size_t doubles)
{
  if(doubles == 0) return;
  gpu::check(fftwpp_gpu_scale((double *) f,s,1,1,doubles,0,0,gpu::stream()),
             "scale");
}

void mismatch(const char *what)
{
  std::cerr << what << std::endl;
  exit(-1);
}

}

fftMPIBase::fftMPIBase(const MPIgroup& group) :
  SlabTranspose(group), app(NULL), fx(NULL), fy(NULL), fz(NULL)
{
}

void fftMPIBase::build(size_t X, size_t Y, size_t Z, bool zp)
{
  d=split3(X,Y,Z,group);
  app=new Application(1,1,multNone,1);
  // x: strided over the y*Z local columns of the X x y x Z layout
  fx=plainPlan(X,*app,std::max<size_t>(d.y,1)*Z);
  // y: contiguous rows (Z == 1) or strided over z inside every x plane
  fy=plainPlan(Y,*app,Z);
  if(zp) fz=plainPlan(Z,*app,1);
}

// Device scratch is allocated on first use, so that the splits and exchange
// tables can be queried (and tested) on a host without a CUDA device.
void fftMPIBase::ready()
{
  devP.ensure(1,n()*sizeof(Complex));
  work.ensure(1,n()*sizeof(Complex));
}

fftMPIBase::~fftMPIBase()
{
  delete fz;
  delete fy;
  delete fx;
  delete app;
}

void fftMPIBase::zpass(int sgn, const void *in, void *out)
{
  transform(fz,sgn,in,out,d.x*d.Y,d.Z,d.x*d.Y*d.Z);
}

void fftMPIBase::ypass(int sgn, const void *in, void *out)
{
  transform(fy,sgn,in,out,d.x,d.Y*d.Z,d.x*d.Y*d.Z);
}

void fftMPIBase::xpass(int sgn, const void *in, void *out)
{
  transform(fx,sgn,in,out,1,0,d.X*d.y*d.Z);
}

// ---------------------------------------------------------------------------

fft2dMPI::fft2dMPI(const split& s, const MPIgroup& group, int sign) :
  fftMPIBase(group), sign(sign)
{
  build(s.X,s.Y,1,false);
}

fft2dMPI::fft2dMPI(size_t X, size_t Y, size_t Z, const MPIgroup& group,
                   int sign) : fftMPIBase(group), sign(sign)
{
  build(X,Y,Z,Z > 1);
}

void fft2dMPI::iForward(Complex *in, Complex *out)
{
  ready();
  if(!out) out=in;
  void *st=gpu::stream();
  if(fz) {
    zpass(sign,in,work.ptr[0]);
    ypass(sign,work.ptr[0],out);
  } else
    ypass(sign,in,out); // rows: every CTA owns whole rows, in place is safe
  transposeBackward(out,work.ptr[0],0,1,st); // x x Y x Z -> X x y x Z
}

void fft2dMPI::ForwardWait(Complex *out)
{
  xpass(sign,work.ptr[0],out);
}

void fft2dMPI::iBackward(Complex *in, Complex *out)
{
  ready();
  if(!out) out=in;
  void *st=gpu::stream();
  xpass(-sign,in,work.ptr[0]);
  transposeForward(work.ptr[0],out,0,1,st); // X x y x Z -> x x Y x Z
}

void fft2dMPI::BackwardWait(Complex *out)
{
  if(fz) {
    ypass(-sign,out,work.ptr[0]);
    zpass(-sign,work.ptr[0],out);
  } else
    ypass(-sign,out,out);
}

void fft2dMPI::Normalize(Complex *f)
{
  scaleDoubles(f,1.0/((double) d.X*d.Y*d.Z),2*d.x*d.Y*d.Z);
}

// ---------------------------------------------------------------------------

rcfft2dMPI::rcfft2dMPI(const split& dr, const split& dc,
                       const MPIgroup& group) : fftMPIBase(group), fr(NULL)
{
  if(dc.X != dr.X || dc.Y != dr.Y/2+1)
    mismatch("rcfft2dMPI: dc must be split(X,Y/2+1) of dr = split(X,Y)");
  setup(dr.X,dr.Y,0);
}

rcfft2dMPI::rcfft2dMPI(size_t X, size_t Y, size_t Z, const MPIgroup& group) :
  fftMPIBase(group), fr(NULL)
{
  setup(X,Y,Z);
}

rcfft3dMPI::rcfft3dMPI(const split3& dr, const split3& dc,
                       const MPIgroup& group) :
  rcfft2dMPI(dr.X,dr.Y,dr.Z,group)
{
  if(dc.X != dr.X || dc.Y != dr.Y || dc.Z != dr.Z/2+1)
    mismatch("rcfft3dMPI: dc must be split3(X,Y,Z/2+1) of dr = split3(X,Y,Z)");
}

void rcfft2dMPI::setup(size_t X, size_t Y, size_t Z)
{
  dims3=Z != 0;
  if(Z == 0) { // 2-D: complex x x Yc, Yc split over the ranks
    rows=1;
    last=Y;
    build(X,Y/2+1,1,false);
    delete fy; // the y pass is the r2c / c2r transform
    fy=NULL;
  } else {
    rows=Y;
    last=Z;
    build(X,Y,Z/2+1,false);
  }
  if(last < 2) mismatch("rcfft: the real dimension must have length >= 2");
  fr=new fftPadReal(last,last,*app,1,1,last,1,true);
}

rcfft2dMPI::~rcfft2dMPI()
{
  delete fr;
}

void rcfft2dMPI::iForward(double *in, Complex *out)
{
  ready();
  void *st=gpu::stream();
  const size_t nr=d.x*rows;
  const size_t lastc=last/2+1;
  void *F=fy ? work.ptr[0] : (void *) out;
  if(nr > 0) // r2c (sign -1) of every row
    gpu::check(fftwpp_gpu_forward(fr->plan(),0,1,1,in,F,nr,last,lastc,st),
               "forward (r2c)");
  if(fy) ypass(-1,F,out);
  transposeBackward(out,work.ptr[0],0,1,st);
}

void rcfft2dMPI::ForwardWait(Complex *out)
{
  xpass(-1,work.ptr[0],out);
}

void rcfft2dMPI::iBackward(Complex *in, double *)
{
  ready();
  void *st=gpu::stream();
  xpass(1,in,work.ptr[0]);
  transposeForward(work.ptr[0],in,0,1,st);
}

void rcfft2dMPI::BackwardWait(Complex *in, double *out)
{
  void *st=gpu::stream();
  const size_t nr=d.x*rows;
  const size_t lastc=last/2+1;
  const void *F=in;
  if(fy) {
    ypass(1,in,work.ptr[0]);
    F=work.ptr[0];
  }
  if(nr > 0) // c2r (sign +1)
    gpu::check(fftwpp_gpu_backward(fr->plan(),0,1,1,F,out,0,1.0,nr,lastc,last,
                                   st),"backward (c2r)");
}

void rcfft2dMPI::Shift(double *f)
{
  void *st=gpu::stream();
  const size_t X=d.X, Y=rows;
  if(X % 2 || (dims3 && Y % 2)) {
    std::cerr << (dims3 ? "Shift is not implemented for odd X or odd Y." :
                  "Shift is not implemented for odd X.") << std::endl;
    exit(1);
  }
  if(d.x == 0) return;
  if(!dims3) { // 2-D: rows with odd global x
    const size_t start=(d.x0+1) % 2;
    if(start < d.x)
      gpu::check(fftwpp_gpu_scale(f+start*last,-1.0,(d.x-start+1)/2,1,last,
                                  2*last,0,st),"shift");
    return;
  }
  // 3-D: rows (i,j) with odd x0+i+j
  for(size_t par=0; par < 2; ++par) { // planes i = par, par+2, ...
    if(par >= d.x) break;
    const size_t ystart=(par+d.x0+1) % 2;
    gpu::check(fftwpp_gpu_scale(f+(par*Y+ystart)*last,-1.0,(d.x-par+1)/2,
                                (Y-ystart+1)/2,last,2*Y*last,2*last,st),
               "shift");
  }
}

// f[i*pitch+k]=0 for i < rows, k < width (words)
void rcfft2dMPI::zeroBox(Complex *f, size_t nrows, size_t width, size_t pitch)
{
  if(nrows == 0 || width == 0) return;
  void *st=gpu::stream();
  ready();
  gpu::check(fftwpp_gpu_memset(work.ptr[0],0,nrows*width*sizeof(Complex),st),
             "memset");
  gpu::check(fftwpp_gpu_memcpy2d(f,pitch*sizeof(Complex),work.ptr[0],
                                 width*sizeof(Complex),width*sizeof(Complex),
                                 nrows,2,st),"zero");
}

void rcfft2dMPI::deNyquist(Complex *f)
{
  const size_t X=d.X;
  if(!dims3) { // 2-D: X x y, y a slice of Y/2+1
    if(X % 2 == 0) zeroBox(f,1,d.y,d.y);
    if(last % 2 == 0 && d.y0+d.y == d.Y && d.y > 0)
      zeroBox(f+d.y-1,X,1,d.y);
    return;
  }
  // 3-D: X x y x Zc
  const size_t yz=d.y*d.Z;
  if(X % 2 == 0) zeroBox(f,1,yz,yz);
  if(rows % 2 == 0 && d.y0 == 0 && d.y > 0) zeroBox(f,X,d.Z,yz);
  if(last % 2 == 0) zeroBox(f+d.Z-1,X*d.y,1,d.Z);
}

void rcfft2dMPI::Normalize(double *f)
{
  scaleDoubles(f,1.0/((double) d.X*rows*last),nreal());
}

}
