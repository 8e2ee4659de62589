/* mpifftw++.h -- distributed 2-D / 3-D FFTs on the convolution path's kernels
 * and exchange (reference mpi/mpifftw++.h:37-585: fft2dMPI, fft3dMPI,
 * rcfft2dMPI, rcfft3dMPI).  Slab decomposition; the exchange is
 * SlabTranspose (NCCL); the 1-D transforms are explicit (q=1) padded-FFT
 * plans: a plan's forward pass is the sign +1 DFT, its adjoint the sign -1
 * DFT, both unnormalised (reference convolve.cc forwardExplicit /
 * backwardExplicit).
 *
 * Layouts as in the reference (mpifftw++.cc:7-80): a transform takes the
 * x x Y [x Z] array (this rank's x rows, all of Y and Z) and leaves the
 * X x y [x Z] array (all X, this rank's y rows).  The arrays are DEVICE
 * pointers with room for n() words.  Forward applies the constructor's sign
 * (default -1), Backward the opposite; Normalize divides by the number of
 * points.  As in the reference, Forward(in,out) may be used in place
 * (out == NULL) and the real-to-complex Backward overwrites its input.
 *
 * Differences from the reference's signatures: the constructors take the
 * process group next to the split (the split classes here do not carry a
 * communicator) and no planning arrays / mpiOptions; in-place real transforms
 * (out aliased to in with padded rows) are not provided.  The non-blocking
 * pairs exist with the reference's names; the exchange is stream-ordered, so
 * iForward enqueues everything up to and including the exchange and
 * ForwardWait the remaining local pass.  Not provided: the pencil (split3
 * xyz) decomposition of fft3dMPI / rcfft3dMPI.  Lengths whose largest
 * prime factor exceeds 64, or that do not fit one CTA's shared memory
 * (> 8192 points), are refused by the plan builder.
 */
#ifndef FFTWPP_B200_MPIFFTWPP_H
#define FFTWPP_B200_MPIFFTWPP_H

#include "mpiconvolve.h"

namespace fftwpp {

// Shared machinery: X x Y x Z complex points (Z == 1: 2-D), split over x
// before and over y after the forward transform.
class fftMPIBase : public SlabTranspose {
public:
  virtual ~fftMPIBase();
  // words (Complex) every complex array must hold
  size_t n() const {return std::max<size_t>(std::max(d.X*d.y,d.x*d.Y)*d.Z,1);}

protected:
  fftMPIBase(const utils::MPIgroup& group);
  Application *app;
  fftBase *fx,*fy,*fz;
  DeviceArrays work;
  void build(size_t X, size_t Y, size_t Z, bool zpass);
  void ready(); // allocates the device scratch on first use
  // complex passes of sign sgn; in != out except for xpass
  void zpass(int sgn, const void *in, void *out);
  void ypass(int sgn, const void *in, void *out);
  void xpass(int sgn, const void *in, void *out);
};

class fft2dMPI : public fftMPIBase {
public:
  fft2dMPI(const utils::split& d, const utils::MPIgroup& group, int sign=-1);

  void iForward(Complex *in, Complex *out=NULL);
  void ForwardWait(Complex *out);
  void Forward(Complex *in, Complex *out=NULL) {
    iForward(in,out);
    ForwardWait(out ? out : in);
  }
  void iBackward(Complex *in, Complex *out=NULL);
  void BackwardWait(Complex *out);
  void Backward(Complex *in, Complex *out=NULL) {
    iBackward(in,out);
    BackwardWait(out ? out : in);
  }
  void Normalize(Complex *f); // on the x x Y layout

protected:
  fft2dMPI(size_t X, size_t Y, size_t Z, const utils::MPIgroup& group,
           int sign);
  int sign;
};

class fft3dMPI : public fft2dMPI {
public:
  fft3dMPI(const utils::split3& d, const utils::MPIgroup& group, int sign=-1) :
    fft2dMPI(d.X,d.Y,d.Z,group,sign) {}
};

// Real-to-complex: real x x Y [x Z] in, complex X x y [x Zc] out; the LAST
// dimension is halved (2-D: Yc = Y/2+1, split over the ranks; 3-D: Zc = Z/2+1,
// y split over the ranks).  dr describes the real array, dc the complex one
// (reference mpifftw++.h:295-585): r2c and the complex passes use sign -1,
// Backward sign +1.
class rcfft2dMPI : public fftMPIBase {
public:
  rcfft2dMPI(const utils::split& dr, const utils::split& dc,
             const utils::MPIgroup& group);
  virtual ~rcfft2dMPI();
  size_t nreal() const {return d.x*rows*last;}

  void iForward(double *in, Complex *out);
  void ForwardWait(Complex *out);
  void Forward(double *in, Complex *out) {
    iForward(in,out);
    ForwardWait(out);
  }
  void iBackward(Complex *in, double *out);
  void BackwardWait(Complex *in, double *out);
  void Backward(Complex *in, double *out) { // overwrites in
    iBackward(in,out);
    BackwardWait(in,out);
  }
  void Normalize(double *f);

  // Transforms with the Fourier origin at the centre of x (and y in 3-D):
  // Shift multiplies the real data by (-1)^x (3-D: (-1)^(x+y)); it needs even
  // X (and Y), as in the reference (mpifftw++.cc:82-101,165-186).
  void Shift(double *f);
  void Forward0(double *in, Complex *out) {
    Shift(in);
    Forward(in,out);
  }
  void Backward0(Complex *in, double *out) {
    Backward(in,out);
    Shift(out);
  }
  // Set the Nyquist modes of even shifted transforms to zero, on the
  // transformed X x y [x Zc] data (mpifftw++.h:345-355,518-538).
  void deNyquist(Complex *f);

protected:
  rcfft2dMPI(size_t X, size_t Y, size_t Z, const utils::MPIgroup& group);
  void zeroBox(Complex *f, size_t rows, size_t width, size_t pitch);
  bool dims3;  // 3-D transform
  size_t rows; // real rows per x plane (2-D: 1, 3-D: Y)
  size_t last; // real row length
  fftBase *fr;
  void setup(size_t X, size_t Y, size_t Z);
};

class rcfft3dMPI : public rcfft2dMPI {
public:
  rcfft3dMPI(const utils::split3& dr, const utils::split3& dc,
             const utils::MPIgroup& group);
};

}

#endif
