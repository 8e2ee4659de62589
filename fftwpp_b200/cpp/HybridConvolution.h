/* HybridConvolution.h -- owning bundles (Application + padded FFTs +
 * convolution) behind the C wrapper API, with the class names, constructor
 * signatures, default padding and public members of the reference's
 * wrappers/HybridConvolution.h:12-190, so that the reference's own
 * wrappers/cfftw++.cc compiles against this directory unchanged
 * (tests/refprogs/Makefile builds it).  lib_fftwpp.so's own C API
 * (cfftwpp.cc) does not use these classes.
 *
 * Default minimal padded sizes: M = A L - A + 1 for the complex bundles,
 * M = 3 ceil(L/2) - 2 (L mod 2) for the Hermitian ones.  Outer dimensions run
 * multNone; the multiplier acts in the innermost dimension.
 */
#ifndef FFTWPP_B200_HYBRIDCONVOLUTION_H
#define FFTWPP_B200_HYBRIDCONVOLUTION_H

#include <vector>

#include "Complex.h"
#include "cfftw++.h"
#include "convolve.h"

using namespace std;
using namespace utils;
using namespace Array;

namespace fftwpp {

namespace hybridbundle {

inline size_t complexPadding(size_t L, size_t A) {return A*L-A+1;}
inline size_t hermitianPadding(size_t L) {
  return 3*utils::ceilquotient(L,2)-2*(L % 2);
}

// Applications and padded FFTs of one bundle, destroyed innermost first.
class Parts {
  std::vector<Application *> apps;
  std::vector<fftBase *> ffts;
public:
  // the Application of the next (inner) dimension: child of the previous one
  Application& app(size_t A, size_t B, multiplier *mult, size_t threads) {
    apps.push_back(apps.empty() ? new Application(A,B,mult,threads) :
                   new Application(A,B,mult,*apps.back()));
    return *apps.back();
  }
  template<class T>
  T *keep(T *fft) {
    ffts.push_back(fft);
    return fft;
  }
  ~Parts() {
    for(size_t i=ffts.size(); i-- > 0;) {
      delete ffts[i];
      delete apps[i];
    }
  }
};

}

class HybridConvolution {
  hybridbundle::Parts parts;
public:
  Convolution *convolve;

  HybridConvolution(size_t L, multiplier mult=multBinary, size_t M=0,
                    size_t A=2, size_t B=1, size_t threads=fftw::maxthreads) {
    if(M == 0) M=hybridbundle::complexPadding(L,A);
    fftPad *fft=parts.keep(new fftPad(L,M,parts.app(A,B,mult,threads)));
    convolve=new Convolution(fft);
  }
  ~HybridConvolution() {delete convolve;}
};

class HybridConvolutionHermitian {
  hybridbundle::Parts parts;
public:
  Convolution *convolve;

  HybridConvolutionHermitian(size_t L, multiplier mult=realMultBinary,
                             size_t M=0, size_t A=2, size_t B=1,
                             size_t threads=fftw::maxthreads) {
    if(M == 0) M=hybridbundle::hermitianPadding(L);
    fftPadHermitian *fft=
      parts.keep(new fftPadHermitian(L,M,parts.app(A,B,mult,threads)));
    convolve=new Convolution(fft);
  }
  ~HybridConvolutionHermitian() {delete convolve;}
};

class HybridConvolution2 {
  hybridbundle::Parts parts;
public:
  Convolution2 *convolve2;

  HybridConvolution2(size_t Lx, size_t Ly, multiplier mult=multBinary,
                     size_t Mx=0, size_t My=0, size_t A=2, size_t B=1,
                     size_t threads=fftw::maxthreads) {
    if(Mx == 0) Mx=hybridbundle::complexPadding(Lx,A);
    if(My == 0) My=hybridbundle::complexPadding(Ly,A);
    fftPad *fftx=
      parts.keep(new fftPad(Lx,Mx,parts.app(A,B,multNone,threads),Ly));
    fftPad *ffty=parts.keep(new fftPad(Ly,My,parts.app(A,B,mult,threads)));
    convolve2=new Convolution2(fftx,ffty);
  }
  ~HybridConvolution2() {delete convolve2;}
};

class HybridConvolutionHermitian2 {
  hybridbundle::Parts parts;
public:
  Convolution2 *convolve2;

  HybridConvolutionHermitian2(size_t Lx, size_t Ly,
                              multiplier mult=realMultBinary, size_t Mx=0,
                              size_t My=0, size_t A=2, size_t B=1,
                              size_t threads=fftw::maxthreads) {
    if(Mx == 0) Mx=hybridbundle::hermitianPadding(Lx);
    if(My == 0) My=hybridbundle::hermitianPadding(Ly);
    const size_t Hy=utils::ceilquotient(Ly,2); // stored modes of y
    fftPadCentered *fftx=parts.keep
      (new fftPadCentered(Lx,Mx,parts.app(A,B,multNone,threads),Hy,Hy));
    fftPadHermitian *ffty=
      parts.keep(new fftPadHermitian(Ly,My,parts.app(A,B,mult,threads)));
    convolve2=new Convolution2(fftx,ffty);
  }
  ~HybridConvolutionHermitian2() {delete convolve2;}
};

class HybridConvolution3 {
  hybridbundle::Parts parts;
public:
  Convolution3 *convolve3;

  HybridConvolution3(size_t Lx, size_t Ly, size_t Lz,
                     multiplier mult=multBinary, size_t Mx=0, size_t My=0,
                     size_t Mz=0, size_t A=2, size_t B=1,
                     size_t threads=fftw::maxthreads) {
    if(Mx == 0) Mx=hybridbundle::complexPadding(Lx,A);
    if(My == 0) My=hybridbundle::complexPadding(Ly,A);
    if(Mz == 0) Mz=hybridbundle::complexPadding(Lz,A);
    fftPad *fftx=
      parts.keep(new fftPad(Lx,Mx,parts.app(A,B,multNone,threads),Ly*Lz));
    fftPad *ffty=
      parts.keep(new fftPad(Ly,My,parts.app(A,B,multNone,threads),Lz));
    fftPad *fftz=parts.keep(new fftPad(Lz,Mz,parts.app(A,B,mult,threads)));
    convolve3=new Convolution3(fftx,ffty,fftz);
  }
  ~HybridConvolution3() {delete convolve3;}
};

class HybridConvolutionHermitian3 {
  hybridbundle::Parts parts;
public:
  Convolution3 *convolve3;

  HybridConvolutionHermitian3(size_t Lx, size_t Ly, size_t Lz,
                              multiplier mult=realMultBinary, size_t Mx=0,
                              size_t My=0, size_t Mz=0, size_t A=2,
                              size_t B=1, size_t threads=fftw::maxthreads) {
    if(Mx == 0) Mx=hybridbundle::hermitianPadding(Lx);
    if(My == 0) My=hybridbundle::hermitianPadding(Ly);
    if(Mz == 0) Mz=hybridbundle::hermitianPadding(Lz);
    const size_t Hz=utils::ceilquotient(Lz,2); // stored modes of z
    fftPadCentered *fftx=parts.keep
      (new fftPadCentered(Lx,Mx,parts.app(A,B,multNone,threads),Ly*Hz));
    fftPadCentered *ffty=parts.keep
      (new fftPadCentered(Ly,My,parts.app(A,B,multNone,threads),Hz));
    fftPadHermitian *fftz=
      parts.keep(new fftPadHermitian(Lz,Mz,parts.app(A,B,mult,threads)));
    convolve3=new Convolution3(fftx,ffty,fftz);
  }
  ~HybridConvolutionHermitian3() {delete convolve3;}
};

}

#endif
