/* convolve.h -- host-side mirror of the reference's hybrid dealiased
 * convolution interface (reference convolve.h), re-implemented for B200.
 *
 * Same class names, constructor signatures, public data members, size
 * queries and error behaviour ("message on cerr + exit", reference
 * convolve.h:226-232) as the reference; together with the companion headers
 * in this directory (Complex.h, fftw++.h, utils.h, align.h, parallel.h,
 * seconds.h, statistics.h, Array.h) the reference's own callers
 * (tests/hybrid*.cc, examples/exampleconv*.cc) compile against it unmodified
 * and link with lib_fftwpp.so -- tests/refprogs/ builds and runs them.  What is different is everything underneath: the objects hold
 * GPU plans (include/fftwpp_gpu.h) instead of FFTW plans, there are no
 * per-routine forward1/forward2/forwardInner variants (one fused kernel family
 * covers them, see csrc/gpu_core.cu), and the residue loop of
 * Convolution{,2,3} collapses into batched launches.
 *
 * Pointer semantics: every entry point that takes data accepts either host
 * pointers (data is staged through device buffers owned by the object; the
 * result is copied back) or device pointers (detected with
 * fftwpp_gpu_is_device_ptr; operated on in place, asynchronously on
 * gpu::stream()).
 *
 * This header contains no CUDA and compiles with a plain C++11 compiler.
 */
#ifndef FFTWPP_B200_CONVOLVE_H
#define FFTWPP_B200_CONVOLVE_H

#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <vector>

// The headers an existing caller gets through the reference's convolve.h
// (reference convolve.h:22-25), so that `#include "convolve.h"` alone keeps
// providing Complex, the aligned allocators, parallel::get_max_threads, the
// timers and the Array classes.
#include "Complex.h"
#include "fftw++.h"
#include "utils.h"
#include "Array.h"

struct fftwpp_gpu_plan;
#include "../../include/fftwpp_gpu.h"

namespace fftwpp {

extern const double twopi;
extern bool showOptTimes;
extern bool showRoutines;

const Complex I(0.0,1.0);

// Smallest 2^a 3^b 5^c 7^d >= m (reference convolve.cc:114-124).
size_t nextfftsize(size_t m);

class fftBase;

// Transformed-index context handed to multipliers (reference convolve.h:48-76).
class Indices {
public:
  fftBase *fft;
  size_t *index;
  size_t size,maxsize;
  size_t r;
  size_t offset;

  Indices() : fft(NULL), index(NULL), size(0), maxsize(0), r(0), offset(0) {}
  void copy(Indices *indices, size_t size0);
  ~Indices() {if(maxsize > 0) delete [] index;}
private:
  Indices(const Indices&);
  Indices& operator=(const Indices&);
};

typedef void multiplier(Complex **F, size_t n, Indices *indices,
                        size_t threads);

// Built-in multipliers (reference convolve.cc:26-110).  Their ADDRESSES select
// the fused device epilogue; the bodies are host implementations used only by
// the unfused custom-multiplier path.
multiplier multNone,multBinary,realMultBinary,multcorrelation;

// Device-side implementation of a user multiplier: same contract as
// `multiplier` (reference convolve.h:78-82), but F[a] are DEVICE pointers to
// the n transformed words of one residue block and the function must only
// ENQUEUE its work (kernel launches) on `stream` (a cudaStream_t).  Without
// one, a user multiplier runs on the host between a GPU forward and a GPU
// backward pass, with a device<->host round trip of the transformed data.
typedef void deviceMultiplier(Complex **F, size_t n, Indices *indices,
                              void *stream);

// Associate a device implementation with the address of a host multiplier:
// every Application constructed with `host` then multiplies on the GPU.
// Passing device=NULL removes the association.
void registerDeviceMultiplier(multiplier *host, deviceMultiplier *device);
deviceMultiplier *deviceMultiplierOf(multiplier *host);

class Application : public ThreadBase {
public:
  size_t A;
  size_t B;
  multiplier *mult;
  bool verbose;
  size_t m;
  size_t D;
  ptrdiff_t I;
  size_t maxthreads;

  void check();

  Application(size_t A, size_t B, multiplier *mult,
              size_t threads=fftw::maxthreads, bool verbose=false,
              size_t m=0, size_t D=0, ptrdiff_t I=-1) :
    ThreadBase(threads), A(A), B(B), mult(mult), verbose(verbose), m(m),
    D(D), I(I), maxthreads(threads) {check();}

  Application(size_t A, size_t B, multiplier *mult, Application &parent,
              size_t m=0, size_t D=0, ptrdiff_t I=-1) :
    ThreadBase(1), A(A), B(B), mult(mult), verbose(parent.verbose), m(m),
    D(D), I(I), maxthreads(parent.maxthreads) {check();}
};

namespace gpu {
// Stream used by all launches of the host classes (default: the legacy
// default stream, which is also torch's default stream).
void *stream();
void setStream(void *cudaStream);
// Abort with the reference's error policy if rc != 0.
void check(int rc, const char *what);
bool isDevice(const void *p);
}

// One call of fft->forward(r): a contiguous range of sub-blocks of the GPU plan.
struct ResidueCall {
  size_t r;      // residue-block argument of forward()/backward()
  size_t sb0;    // first sub-block
  size_t nsb;    // number of sub-blocks
  size_t rows;   // output rows (FFT outputs per column) produced by the call
  size_t row0;   // first row in the all-residues layout
};

class fftBase : public ThreadBase {
public:
  size_t L; // number of unpadded data values
  size_t M; // minimum number of padded data values (becomes m*q)
  size_t C; // number of FFTs to compute in parallel
  size_t S; // stride between successive elements
  size_t m;
  size_t p;
  size_t q;
  size_t n;  // number of residues
  size_t R;  // number of residue blocks
  size_t dr; // r increment
  size_t D;  // number of residues stored in F at a time
  size_t D0; // remainder
  size_t Cm,Sm;
  size_t l;  // block size of a single FFT
  size_t b;  // total block size, including stride
  bool inplace;
  Application app;
  bool centered;
  bool overwrite; // always false: the GPU path never overwrites its input

  enum Kind {COMPLEX=0, CENTERED=1, HERMITIAN=2, REAL=3};

  static void parameters(size_t L, size_t M, size_t m, bool centered,
                         size_t &p, size_t& n, size_t& q);

  virtual ~fftBase();

  void invalid();

  // No-op: implicit padding happens inside the kernels.
  void pad(Complex *) {}

  // Residue-block transforms with the reference's call shape
  // (convolve.h:262-268).  f, F: host or device pointers.  W is ignored.
  void forward(Complex *f, Complex *F, size_t r=0, Complex *W=NULL);
  void backward(Complex *F, Complex *f, size_t r=0, Complex *W=NULL);

  virtual size_t index(size_t r, size_t i);

  size_t normalization() {return M;}
  size_t paddedSize() {return m*q;}

  virtual bool conjugates() {
    return D > 1 && (p <= 2 || (centered && p % 2 == 0));
  }
  virtual size_t residueBlocks() {
    return conjugates() ? utils::ceilquotient(n,2) : n;
  }
  size_t Dr() {return conjugates() ? D/2 : D;}
  virtual size_t increment(size_t r) {
    return r > 0 ? dr : (conjugates() ? utils::ceilquotient(D0,2) : D0);
  }
  size_t nloops();
  bool loop2() {return nloops() == 2 && app.A > app.B && !overwrite;}

  virtual size_t inputLength() {return L;}
  virtual size_t wordSize() {return 2;}
  virtual size_t doubles() {return wordSize()*S*inputLength();}
  virtual size_t outputSize() {return b*D;}
  virtual size_t blocksize(size_t) {return l;}
  virtual size_t noutputs(size_t r) {return blocksize(r)*(r == 0 ? D0 : D);}
  virtual size_t span(size_t r) {return S*noutputs(r);}
  size_t workSizeV() {
    return nloops() == 1 || loop2() ? 0 : utils::ceilquotient(doubles(),2);
  }
  virtual size_t workSizeW() {return inplace ? 0 : outputSize();}
  size_t repad() {return !inplace && L < m;}

  // Median time in ns of one convolveRaw on the device (reference time()).
  double time();
  double report();

  // ---- GPU-side description (not in the reference) ----
  virtual Kind kind()=0;
  // The GPU plan is created on first use so that the bookkeeping above can be
  // queried (and tested) on a host without a CUDA device.
  fftwpp_gpu_plan *plan();
  // Profiling tag of this pass (1 = x, 2 = y, 3 = z), see fftwpp_gpu_profile_*.
  void setTag(int tag);
  int tag() {return gputag;}
  const std::vector<ResidueCall>& calls() {return callTable;}
  // Rows (FFT outputs per column) when all residues are produced at once.
  size_t allRows() {return totalRows;}
  // Words (Complex; doubles for Hermitian) of an all-residues output buffer.
  size_t allSize() {return totalRows*S;}
  const ResidueCall& call(size_t r);

  // Two-stage ("inner", p > 2) execution of the fused 1-D convolution for
  // large power-of-two transforms (reference forwardInner/backwardInner,
  // convolve.cc:1227-1466,1765-1965): stage A = length-p pass over the m
  // interleaved columns + outer twiddle, stage B = q explicit length-m rows.
  bool innerFast() {return twoStage;}
  fftwpp_gpu_plan *innerA();
  fftwpp_gpu_plan *innerB();

protected:
  static bool innerEligible(Kind kind, size_t L, size_t m, size_t p, size_t C,
                            size_t S);
  fftwpp_gpu_plan *gpuplan;
  int gputag;
  std::vector<struct SubBlockHost> *subHost;
  std::vector<ResidueCall> callTable;
  size_t totalRows;
  void *devIn,*devOut; // staging for host-pointer forward()/backward()
  bool twoStage;
  bool forcedCtor; // built by a forced (m,D,I) constructor
  fftwpp_gpu_plan *planA,*planB;

  fftBase(size_t L, size_t M, Application& app, size_t C, size_t S,
          bool centered);
  fftBase(size_t L, size_t M, Application& app, size_t C, size_t S, size_t m,
          size_t D, bool inplace, bool centered);

  void checkParameters();
  void common();
  // Heuristic replacement of the reference's timing optimizer
  // (convolve.cc:412-470): fills m,D,inplace honouring app.m/app.D/app.I.
  void choose(bool Explicit);
  virtual bool valid(size_t m, size_t p, size_t q, size_t n, size_t D,
                     size_t S)=0;
  void report(const char *name);
  void buildPlan(const std::vector<struct SubBlockHost>& sub);
  void stage(Complex *&devF, Complex *&devf, Complex *f, Complex *F,
             bool toDeviceF, bool toDevicef);
};

class fftPad : public fftBase {
public:
  static bool valid0(size_t m, size_t p, size_t q, size_t n, size_t D,
                     size_t S) {
    if(q == 1) return D == 1;
    return D == 1 || (S == 1 && ((D < n && D % 2 == 0) || D == n));
  }

  fftPad(size_t L, size_t M, Application& app, size_t C=1, size_t S=0,
         bool Explicit=false);
  fftPad(size_t L, size_t M, Application &app, size_t C, size_t S, size_t m,
         size_t D, bool inplace);
  Kind kind() {return centered ? CENTERED : COMPLEX;}

protected:
  struct Deferred {};
  fftPad(size_t L, size_t M, Application &app, size_t C, size_t S, Deferred);
  fftPad(size_t L, size_t M, Application &app, size_t C, size_t S, size_t m,
         size_t D, bool inplace, Deferred);
  bool valid(size_t m, size_t p, size_t q, size_t n, size_t D, size_t S) {
    return valid0(m,p,q,n,D,S);
  }
  void init();
};

class fftPadCentered : public fftPad {
public:
  fftPadCentered(size_t L, size_t M, Application& app, size_t C=1, size_t S=0,
                 bool Explicit=false);
  fftPadCentered(size_t L, size_t M, Application &app, size_t C, size_t S,
                 size_t m, size_t D, bool inplace);
  bool conjugates() {return D > 1 && (p == 1 || p % 2 == 0);}
protected:
  bool valid(size_t m, size_t p, size_t q, size_t n, size_t D, size_t S) {
    return (q == 1 || p % 2 == 0) && valid0(m,p,q,n,D,S);
  }
};

class fftPadHermitian : public fftBase {
  size_t e;
  size_t B; // work block size
public:
  fftPadHermitian(size_t L, size_t M, Application& app, size_t C=1,
                  bool Explicit=false);
  fftPadHermitian(size_t L, size_t M, Application &app, size_t C, size_t m,
                  size_t D, bool inplace);
  Kind kind() {return HERMITIAN;}

  size_t inputLength() {return utils::ceilquotient(L,2);}
  size_t blocksize(size_t) {return m*(q == 1 ? 1 : p/2);}
  size_t noutputs(size_t) {return blocksize(0);}
  size_t span(size_t) {return 2*b;}
  size_t workSizeW() {return inplace ? 0 : B*D;}
protected:
  bool valid(size_t m, size_t p, size_t q, size_t n, size_t D, size_t C) {
    return (D == 1 && q == 1) || (D == 2 && p % 2 == 0 && (p == 2 || C == 1));
  }
  void init();
};

class fftPadReal : public fftBase {
  size_t e;
public:
  fftPadReal(size_t L, size_t M, Application& app, size_t C=1, size_t S=0,
             bool Explicit=false);
  fftPadReal(size_t L, size_t M, Application &app, size_t C, size_t S,
             size_t m, size_t D, bool inplace);
  Kind kind() {return REAL;}

  size_t wordSize() {return 1;}
  size_t outputSize() {
    if(n == 2) return p > 2 ? (p/2+1)*m*S : e*S;
    return b*D;
  }
  bool conjugates() {return false;}
  size_t residueBlocks() {return utils::ceilquotient(n+1,2);}
  size_t increment(size_t r) {return r > 1 ? D : r == 1 ? D0 : 1;}
  size_t blocksize(size_t r) {
    if(r == 0) return p > 2 ? (p % 2 ? (p/2+1)*m : (p/2)*m+e-1) : e;
    if(2*r == n) return p > 2 ? (p/2)*m : e-1;
    return l;
  }
  size_t noutputs(size_t r) {
    if(r == 0) return blocksize(0);
    return blocksize(r)*(2*r == n ? 1 : r == 1 ? D0 : D);
  }
  size_t index(size_t r, size_t i);
protected:
  bool valid(size_t m, size_t p, size_t q, size_t n, size_t D, size_t S) {
    return (n % 2 == 1 || (p % 2 == 0 || p <= 2)) && (q % 2 == 1 || m % 2 == 0)
      && (D == 1 || (S == 1 && ((D < (n-1)/2 && D % 2 == 0) || D == (n-1)/2)));
  }
  void init();
};

// Owned device (or staging) storage of a convolution object.
struct DeviceArrays {
  std::vector<void *> ptr;
  size_t bytesEach;
  DeviceArrays() : bytesEach(0) {}
  void ensure(size_t count, size_t bytes);
  void release();
  ~DeviceArrays() {release();}
};

class Convolution : public ThreadBase {
public:
  fftBase *fft;
  size_t L;
  size_t A;
  size_t B;
  multiplier *mult;
  double scale;
  Indices indices;

  // F, W, V are accepted for source compatibility and ignored: padded data
  // lives in shared memory only.
  Convolution(fftBase *fft, Complex **F=NULL, Complex *W=NULL,
              Complex *V=NULL);
  virtual ~Convolution();

  double normalization() {return fft->normalization();}
  size_t increment(size_t r) {return fft->increment(r);}

  void convolveRaw(Complex **f);
  void convolveRaw(Complex **f, Indices *indices);
  void convolveRaw(Complex **f, size_t offset);
  void convolveRaw(Complex **f, size_t offset, Indices *indices);
  void convolveRaw(double **f) {convolveRaw((Complex **) f);}

  void convolve(Complex **f);
  void convolve(Complex **f, size_t offset);
  void convolve(double **f) {convolve((Complex **) f);}

  void normalize(Complex **h, size_t offset=0);

  // Batched device entry used by Convolution2/3 (replaces the OpenMP loop over
  // x rows, reference convolve.h:1434-1445): nrows independent rows, row i of
  // array a at f[a]+offset+i*rowstride (input words).  Device pointers only.
  void convolveRows(Complex **f, size_t offset, size_t nrows,
                    size_t rowstride, double scale);

  // Transformed outer indices of the rows of the next convolveRows() batch,
  // `dims` entries per row (row-major): what the reference's per-row loops
  // store in indices.index[d] before each inner convolveRaw
  // (reference convolve.h:1442,1759).  Only consulted for user multipliers.
  void setRowIndices(const std::vector<size_t>& table, size_t dims) {
    rowIndex=table; rowIndexDims=dims;
  }
  bool customMultiplier() const {return multId < 0;}

protected:
  int multId;
  std::vector<size_t> rowIndex;
  size_t rowIndexDims;
  DeviceArrays dev;
  DeviceArrays devT; // stage-A output of two-stage transforms
  void run(Complex **f, size_t offset, double scale);
  void runCustom(Complex **f, size_t offset, size_t nrows, size_t rowstride,
                 double scale);
};

// Enforce Hermitian symmetry on host data (reference convolve.h:1168-1267).
inline void HermitianSymmetrize(Complex *f)
{
  f[0]=Complex(f[0].real(),0.0);
}
void HermitianSymmetrizeX(size_t Hx, size_t Hy, size_t x0, Complex *f,
                          size_t Sx, size_t threads=fftw::maxthreads);
inline void HermitianSymmetrizeX(size_t Hx, size_t Hy, size_t x0, Complex *f)
{
  HermitianSymmetrizeX(Hx,Hy,x0,f,Hy);
}
void HermitianSymmetrizeXY(size_t Hx, size_t Hy, size_t Hz, size_t x0,
                           size_t y0, Complex *f, size_t Sx, size_t Sy,
                           size_t threads=fftw::maxthreads);
inline void HermitianSymmetrizeXY(size_t Hx, size_t Hy, size_t Hz, size_t x0,
                                  size_t y0, Complex *f)
{
  size_t Ly=y0+Hy;
  HermitianSymmetrizeXY(Hx,Hy,Hz,x0,y0,f,Ly*Hz,Hz);
}

class Convolution2 : public ThreadBase {
public:
  fftBase *fftx,*ffty;
  Convolution **convolvey;
  size_t Lx,Ly; // x,y dimensions of input data
  size_t Sx;    // x stride
  size_t A;
  size_t B;
  multiplier *mult;
  double scale;
  Indices indices;

  Convolution2(fftBase *fftx, fftBase *ffty, Complex **F=NULL,
               Complex *W=NULL, Complex *V=NULL);
  virtual ~Convolution2();

  double normalization() {
    return fftx->normalization()*convolvey[0]->normalization();
  }

  virtual size_t blocksizex(size_t rx) {return fftx->blocksize(rx);}
  virtual size_t stridex() {return Sx;}
  virtual size_t indexBase() {return 0;}
  virtual size_t inputLengthy() {return ffty->inputLength();}

  void convolveRaw(Complex **f, size_t offset=0, Indices *indices=NULL);
  void convolveRaw(double **f, size_t offset=0, Indices *indices=NULL) {
    convolveRaw((Complex **) f,offset,indices);
  }
  void convolve(Complex **f, size_t offset=0);
  void convolve(double **f, size_t offset=0) {convolve((Complex **) f,offset);}
  void normalize(Complex **h, size_t offset=0);

  // Batched device entry used by Convolution3: nplanes independent x-y planes,
  // plane i of array a at f[a]+offset+i*planestride.  Device pointers only.
  void convolvePlanes(Complex **f, size_t offset, size_t nplanes,
                      size_t planestride, double scale);

  // Number of planes processed per batch by convolvePlanes (0 = all).  Small
  // batches keep the y/z intermediates resident in the 126 MB L2.
  size_t planeChunk;

  // Fused exchange (distributed runs): when set, the backward strided pass of
  // output b stores row j of plane i at (Complex *) outBase[b][j] +
  // i*outStride[b][j] + column -- possibly in a peer GPU's memory -- instead
  // of back into F[b] (device arrays; see fftwpp_gpu_backward_mapped).
  std::vector<const uint64_t *> outBase;
  std::vector<const int64_t *> outStride;
  // The same destinations as per-owner row ranges (fftwpp_gpu_backward_dests:
  // bulk tensor stores issued by the TMA unit); tried first.
  std::vector<std::vector<fftwpp_gpu_dest> > outDests;

  // Transformed index (dimension above this object's) of each plane of the
  // next convolvePlanes() batch; set by Convolution3 for user multipliers
  // (reference convolve.h:1759: cyz->indices.index[1]=fftx->index(rx,i+base)).
  std::vector<size_t> planeIndex;

protected:
  DeviceArrays dev;   // staging of host inputs
  DeviceArrays devF;  // x-transformed data, all residues
  size_t FxBytes;
  void run(Complex **f, size_t offset, double scale);
};

class Convolution3 : public ThreadBase {
public:
  fftBase *fftx,*ffty,*fftz;
  Convolution **convolvez;
  Convolution2 **convolveyz;
  size_t Lx,Ly,Lz; // x,y,z dimensions of input data
  size_t Sx,Sy;    // x stride, y stride
  size_t A;
  size_t B;
  multiplier *mult;
  double scale;
  Indices indices;

  Convolution3(fftBase *fftx, fftBase *ffty, fftBase *fftz, Complex **F=NULL,
               Complex *W=NULL, Complex *V=NULL, bool mpi=false);
  virtual ~Convolution3();

  double normalization() {
    return fftx->normalization()*convolveyz[0]->normalization();
  }
  bool contiguous() {return Sy == Lz || fftx->wordSize() != 2;}
  void checkStrides();

  virtual size_t blocksizex(size_t rx) {return fftx->blocksize(rx);}
  virtual size_t stridex() {return Sx;}
  virtual size_t indexBase() {return 0;}
  virtual size_t inputLengthy() {return ffty->inputLength();}
  virtual size_t inputLengthz() {return fftz->inputLength();}

  virtual void convolveRaw(Complex **f, size_t offset=0,
                           Indices *indices=NULL);
  void convolveRaw(double **f, size_t offset=0, Indices *indices=NULL) {
    convolveRaw((Complex **) f,offset,indices);
  }
  virtual void convolve(Complex **f, size_t offset=0);
  void convolve(double **f, size_t offset=0) {convolve((Complex **) f,offset);}
  void normalize(Complex **h, size_t offset=0);

protected:
  DeviceArrays dev;
  DeviceArrays devF;
  void run(Complex **f, size_t offset, double scale);
};

} // namespace fftwpp

#endif
