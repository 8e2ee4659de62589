// convolve.cc -- host classes of the B200 hybrid dealiased convolution
// (see convolve.h).  Plans, partitions and residue bookkeeping live here; all
// arithmetic on data happens in CUDA kernels reached through the thin C ABI
// (include/fftwpp_gpu.h).  There is no CPU compute path.
//
// Residue bookkeeping (p,q,n,D,D0,dr,R,l,b, index(), increment(), sizes)
// restates reference convolve.cc:403-410,493-509,511-719,4309-4421,5449-5637
// and convolve.h:297-455,786-802,925-979 so that callers that walk these
// accessors (tests/hybrid*.cc) see reference-consistent values.

#include "convolve.h"
#include "../../include/fftwpp_gpu.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

using namespace utils;

namespace utils {
// Must be a power of two and at least sizeof(Complex) (reference parallel.cc:18)
size_t ALIGNMENT=2*sizeof(Complex);
}

// Host-thread globals of the reference's parallel.cc (:24,28,65-74).  The GPU
// path does no host-threaded arithmetic, so there is no threshold to measure.
size_t threshold=SIZE_MAX;
namespace parallel {
size_t lastThreads=SIZE_MAX;
void Threshold(size_t threads) {lastThreads=threads;}
}

namespace fftwpp {

const double twopi=2.0*M_PI;
bool showOptTimes=false;
bool showRoutines=false;

size_t fftw::maxthreads=1;
size_t fftw::effort=0;

// ---------------------------------------------------------------------------
// gpu plumbing
// ---------------------------------------------------------------------------

namespace gpu {

static void *currentStream=NULL;

void *stream() {return currentStream;}
void setStream(void *s) {currentStream=s;}

void check(int rc, const char *what)
{
  if(rc == 0) return;
  std::cerr << "fftwpp-b200: " << what << " failed (" << rc << "): "
            << fftwpp_gpu_last_error() << std::endl;
  exit(-1);
}

bool isDevice(const void *p)
{
  return fftwpp_gpu_is_device_ptr(p) > 0;
}

} // namespace gpu

void DeviceArrays::ensure(size_t count, size_t bytes)
{
  if(ptr.size() >= count && bytesEach >= bytes) return;
  release();
  ptr.assign(count,NULL);
  bytesEach=bytes;
  for(size_t i=0; i < count; ++i)
    gpu::check(fftwpp_gpu_malloc(&ptr[i],bytes),"device allocation");
}

void DeviceArrays::release()
{
  for(size_t i=0; i < ptr.size(); ++i)
    if(ptr[i]) fftwpp_gpu_free(ptr[i]);
  ptr.clear();
  bytesEach=0;
}

// ---------------------------------------------------------------------------
// multipliers (host bodies; the fused device epilogues are selected by address)
// ---------------------------------------------------------------------------

void multNone(Complex **, size_t, Indices *, size_t) {}

void multBinary(Complex **F, size_t n, Indices *, size_t)
{
  Complex *F0=F[0], *F1=F[1];
  for(size_t j=0; j < n; ++j) F0[j] *= F1[j];
}

void realMultBinary(Complex **F, size_t n, Indices *, size_t)
{
  double *F0=(double *) F[0], *F1=(double *) F[1];
  for(size_t j=0; j < n; ++j) F0[j] *= F1[j];
}

void multcorrelation(Complex **F, size_t n, Indices *, size_t)
{
  Complex *F0=F[0], *F1=F[1];
  for(size_t j=0; j < n; ++j) F0[j] *= conj(F1[j]);
}

// max(A,B) the fused convolution kernels accept (MAXARRAYS of the CUDA layer)
static const size_t MAXFUSED=8;

static int multiplierId(multiplier *mult)
{
  if(mult == multNone) return FFTWPP_MULT_NONE;
  if(mult == multBinary) return FFTWPP_MULT_BINARY;
  if(mult == realMultBinary) return FFTWPP_MULT_REALBINARY;
  if(mult == multcorrelation) return FFTWPP_MULT_CORRELATION;
  return -1; // custom host multiplier: unfused path
}

namespace {
std::vector<std::pair<multiplier *,deviceMultiplier *> >& deviceMultipliers()
{
  static std::vector<std::pair<multiplier *,deviceMultiplier *> > table;
  return table;
}
}

void registerDeviceMultiplier(multiplier *host, deviceMultiplier *device)
{
  auto& t=deviceMultipliers();
  for(size_t i=0; i < t.size(); ++i)
    if(t[i].first == host) {
      if(device) t[i].second=device;
      else t.erase(t.begin()+i);
      return;
    }
  if(device) t.push_back(std::make_pair(host,device));
}

deviceMultiplier *deviceMultiplierOf(multiplier *host)
{
  auto& t=deviceMultipliers();
  for(size_t i=0; i < t.size(); ++i)
    if(t[i].first == host) return t[i].second;
  return NULL;
}

void Indices::copy(Indices *indices, size_t size0)
{
  size=indices ? indices->size : size0;
  if(size > maxsize) {
    if(maxsize > 0) delete [] index;
    index=new size_t[size];
    maxsize=size;
    for(size_t d=0; d < size; ++d) index[d]=0;
  }
  if(indices)
    for(size_t d=1; d < size; ++d)
      index[d]=indices->index[d];
}

void Application::check()
{
  if(m == 1) {
    std::cerr << std::endl
              << "WARNING: m=1 changed to m=0 to force optimization."
              << std::endl;
    m=0;
  }
}

size_t nextfftsize(size_t m)
{
  size_t N=ceilpow2(m);
  if(m == N) return m;
  for(size_t a=1; a < N; a *= 7)
    for(size_t b=a; b < N; b *= 5)
      for(size_t c=b; c < N; c *= 3)
        N=std::min(N,c*ceilpow2(ceilquotient(m,c)));
  return N;
}

static bool ispow2(size_t m) {return m > 0 && (m & (m-1)) == 0;}

// ---------------------------------------------------------------------------
// fftBase
// ---------------------------------------------------------------------------

struct SubBlockHost {
  fftwpp_gpu_subblock s;
};

fftBase::fftBase(size_t L, size_t M, Application& app, size_t C, size_t S,
                 bool centered) :
  ThreadBase(app.threads), L(L), M(M), C(C), S(S == 0 ? C : S), m(0), p(0),
  q(0), n(0), R(0), dr(0), D(0), D0(0), Cm(0), Sm(0), l(0), b(0),
  inplace(false), app(app), centered(centered), overwrite(false),
  gpuplan(NULL), gputag(0), subHost(NULL), totalRows(0), devIn(NULL), devOut(NULL),
  twoStage(false), planA(NULL), planB(NULL)
{
  forcedCtor=false;
  checkParameters();
}

fftBase::fftBase(size_t L, size_t M, Application& app, size_t C, size_t S,
                 size_t m, size_t D, bool inplace, bool centered) :
  ThreadBase(app.threads), L(L), M(M), C(C), S(S == 0 ? C : S), m(m), p(0),
  q(0), n(0), R(0), dr(0), D(D), D0(0), Cm(0), Sm(0), l(0), b(0),
  inplace(inplace), app(app), centered(centered), overwrite(false),
  gpuplan(NULL), gputag(0), subHost(NULL), totalRows(0), devIn(NULL), devOut(NULL),
  twoStage(false), planA(NULL), planB(NULL)
{
  checkParameters();
  this->app.D=D;
  forcedCtor=true; // no optimizer scan => no parameter report (reference
                   // convolve.cc:451-469 prints from OptBase::scan only)
}

fftBase::~fftBase()
{
  if(planA) fftwpp_gpu_plan_destroy(planA);
  if(planB) fftwpp_gpu_plan_destroy(planB);
  if(gpuplan) fftwpp_gpu_plan_destroy(gpuplan);
  delete subHost;
  if(devIn) fftwpp_gpu_free(devIn);
  if(devOut) fftwpp_gpu_free(devOut);
}

void fftBase::invalid()
{
  std::cerr << "Invalid parameters: " << std::endl
            << "m=" << m << " p=" << p << " q=" << q
            << " n=" << n << " " << " D=" << D << " S=" << S
            << std::endl;
  exit(-1);
}

void fftBase::checkParameters()
{
  if(L > M) {
    std::cerr << "L=" << L << " is greater than M=" << M << "." << std::endl;
    exit(-1);
  }
  if(S < C) {
    std::cerr << "stride S cannot be less than count C" << std::endl;
    exit(-1);
  }
}

void fftBase::parameters(size_t L, size_t M, size_t m, bool centered,
                         size_t &p, size_t& n, size_t& q)
{
  p=ceilquotient(L,m);
  // effective number of length-m blocks the data is folded into
  size_t P=((centered && p % 2 == 0) || p == 2) ? p/2 : p;
  n=ceilquotient(M,P*m);
  q=P*n;
}

void fftBase::common()
{
  parameters(L,M,m,centered,p,n,q);
  if(q*m < M) {
    std::cerr << "Invalid parameters: " << std::endl
              << " q=" << q << " m=" << m << " M=" << M << std::endl;
    exit(-1);
  }
  Cm=C*m;
  Sm=S*m;
  M=m*q;
  overwrite=false;
}

size_t fftBase::nloops()
{
  size_t count=0;
  for(size_t r=0; r < R; r += increment(r))
    ++count;
  return count;
}

// Position i of the output of forward(r) holds the transform at this global
// index of the padded FFT (restates reference convolve.h:297-326).
size_t fftBase::index(size_t r, size_t i)
{
  if(q == 1) return i;
  const size_t P=ceilquotient(p,2);
  size_t s=i % m;
  size_t u;
  const bool paired=D > 1 && ((centered && p % 2 == 0) || p <= 2);
  if(paired) {
    const size_t Pm=P*m;
    u=(i/m) % P;
    const size_t lead=(r == 0 && i >= Pm && D0 % 2 == 1) ? 1 : 0;
    const size_t pair=(i+Pm*lead)/(2*Pm);
    r += pair;
    if(i/Pm-2*pair+lead == 1) { // conjugate partner of the pair
      if((!centered && p == 2) || (r > 0 && u == 0))
        s=s > 0 ? s-1 : m-1;
      if(r == 0)
        r=n/2;
      else {
        r=n-r;
        u=u > 0 ? u-1 : P-1;
      }
    }
  } else {
    u=(i/m) % p;
    r += i/(p*m);
  }
  return q*s+n*u+r;
}

const ResidueCall& fftBase::call(size_t r)
{
  for(size_t i=0; i < callTable.size(); ++i)
    if(callTable[i].r == r) return callTable[i];
  std::cerr << "Invalid residue block r=" << r << " (R=" << R << ")"
            << std::endl;
  exit(-1);
}

void fftBase::report(const char *name)
{
  if(!app.verbose || forcedCtor) return;
  size_t mpL=m*p-L;
  std::cout << std::endl << "Optimal padding: ";
  if(p == q) std::cout << "Explicit" << std::endl;
  else if(mpL > 0) std::cout << "Hybrid" << std::endl;
  else std::cout << "Implicit" << std::endl;
  std::cout << "m=" << m << std::endl;
  std::cout << "p=" << p << std::endl;
  std::cout << "q=" << q << std::endl;
  std::cout << "C=" << C << std::endl;
  std::cout << "S=" << S << std::endl;
  std::cout << "D=" << D << std::endl;
  std::cout << "I=" << inplace << std::endl;
  std::cout << "threads=" << app.threads << std::endl;
  std::cout << "Padding: " << mpL << std::endl;
  if(showRoutines)
    std::cout << "Forwards Routine: " << name << "::gpuForward" << std::endl
              << "Backwards Routine: " << name << "::gpuBackward" << std::endl;
}

// Deterministic chooser standing in for the reference's timing optimizer
// (convolve.cc:239-470).  Candidates follow the reference's enumeration
// (SURVEY Appendix B); the timing is replaced by a cost model of the GPU
// kernels: work ~ N (log2 m + 2 + p), non-power-of-two m pays the generic
// mixed-radix kernel, and a candidate must fit the shared-memory tile.
// FFTWPP_NO_LONG_ROWS=1: rows of 4096 < L <= 8192 take the two-stage (inner)
// path instead of the tensor-memory row kernel
static bool longRows()
{
  const char *s=getenv("FFTWPP_NO_LONG_ROWS");
  return !(s && *s && *s != '0');
}

// Rows of 8192 points exist only as the fused A=2, B=1 tensor-memory kernel
// (fast_conv_rows_long): stage B of a two-stage transform may use them for the
// built-in binary multipliers only.
static bool longStageB(const Application& app)
{
  return app.A == 2 && app.B == 1 &&
    (app.mult == multBinary || app.mult == multcorrelation);
}

void fftBase::choose(bool Explicit)
{
  const bool mForced=app.m >= 1;
  const bool DForced=app.D > 0;
  std::vector<size_t> cand;
  if(Explicit) {
    if(mForced && app.m >= M) cand.push_back(app.m);
    else {
      cand.push_back(nextfftsize(M));
      if(C == 1) cand.push_back(ceilpow2(M));
    }
  } else if(mForced) {
    cand.push_back(app.m);
  } else {
    size_t H=ceilquotient(L,2);
    cand.push_back(ceilpow2(L));
    cand.push_back(nextfftsize(L));
    cand.push_back(ceilpow2(H));
    cand.push_back(nextfftsize(H));
    cand.push_back(nextfftsize(ceilquotient(M,2)));
    cand.push_back(nextfftsize(M));
    if(C == 1) cand.push_back(ceilpow2(M));
    for(size_t mi=16; mi < H && mi <= 8192; mi *= 2)
      cand.push_back(mi);
  }

  const size_t smemBytes=200*1024;
  const size_t Lin=inputLength();
  const size_t word=wordSize()*sizeof(double);
  double best=1e300;
  bool found=false;
  size_t bestm=0, bestD=1;
  for(size_t ic=0; ic < cand.size(); ++ic) {
    size_t mc=cand[ic];
    if(mc == 0) continue;
    size_t pc,nc,qc;
    parameters(L,Explicit ? mc : M,mc,centered,pc,nc,qc);
    if(mc*qc < (Explicit ? mc : M)) continue;
    size_t Dc=DForced ? app.D : (kind() == HERMITIAN && qc > 1 ? 2 : 1);
    if(!valid(mc,pc,qc,nc,Dc,kind() == HERMITIAN ? C : S)) continue;
    // the chooser does not pick the inner (p > 2) layouts of the Hermitian
    // and real classes on its own; forced values (Application.m) reach them
    if(!mForced && qc > 1 && pc > 2 && (kind() == HERMITIAN || kind() == REAL))
      continue;
    size_t lane=(C == 1 ? (app.A+app.B)*Lin*word+
                 std::max(app.A,app.B)*mc*sizeof(Complex) :
                 Lin*word+mc*sizeof(Complex));
    bool inner=qc > 1 && innerEligible(kind(),L,mc,pc,C,S) &&
      (mc <= 4096 || longStageB(app));
    // the fused register kernels (fast_kernels.cu) hold a whole power-of-two
    // row of up to 4096 points (2048 with two input terms) on chip
    // regardless of the generic kernels' tile estimate; p=1, q=2 rows of 8192
    // points run on the tensor-memory kernel (fast_conv_rows_long)
    bool fusedRow=C == 1 && kind() == COMPLEX && app.A == 2 && app.B == 1 &&
      (app.mult == multBinary || app.mult == multcorrelation) &&
      ispow2(mc) && mc >= 16 &&
      (pc == 1 ? (mc <= 4096 || (mc == 8192 && qc == 2 && longRows())) :
       pc == 2 && mc <= 2048);
    if(!mForced && !inner && !fusedRow && lane > smemBytes) continue;
    double N=(double) mc*qc;
    double cost;
    if(inner) { // two global passes + one fused pass
      double lm=log2((double) mc), lp=log2((double) pc);
      // measured on B200 (L=2^13 rows and L=2^20): flat optimum around
      // m = 4p..8p (the fused stage prefers long rows, the strided stage
      // short transforms); rows of m >= 4096 with the tensor-memory row
      // kernel serving stage B (A=2, B=1, binary multipliers) are cheaper
      // (profiles/long_rows_r02.jsonl: L=2^20 m=4096 0.128 ms, m=2048 0.150;
      // L=2^22 m=8192 0.523, m=4096 0.575, m=2048 0.679)
      cost=N*(lm+lp+8.0+0.25*fabs(lm-lp-2.5));
      if(mc >= 4096 && longStageB(app)) cost *= mc == 8192 ? 0.87 : 0.88;
    } else {
      cost=N*(log2((double) mc)+2.0+pc)*(ispow2(mc) ? 1.0 : 2.5);
      if(pc > 2) cost *= 1.0+0.25*pc;
    }
    if(!found || cost < best) {
      found=true;
      best=cost;
      bestm=mc;
      bestD=Dc;
    }
  }
  if(!found) {
    std::cerr << "Optimizer found no valid cases with specified parameters."
              << std::endl;
    std::cerr << "Using explicit routines with m=" << M
              << ", D=1, and I=0 instead." << std::endl << std::endl;
    m=M;
    D=1;
    inplace=false;
    return;
  }
  m=bestm;
  D=bestD;
  inplace=app.I == -1 ? true : app.I != 0;
}

void fftBase::buildPlan(const std::vector<SubBlockHost>& sub)
{
  delete subHost;
  subHost=new std::vector<SubBlockHost>(sub);
}

fftwpp_gpu_plan *fftBase::plan()
{
  if(gpuplan) return gpuplan;
  std::vector<fftwpp_gpu_subblock> s(subHost->size());
  for(size_t i=0; i < s.size(); ++i) s[i]=(*subHost)[i].s;
  fftwpp_gpu_pad_desc d;
  memset(&d,0,sizeof(d));
  d.kind=(int) kind();
  d.L=L;
  d.Lin=inputLength();
  d.N=m*q;
  d.m=m;
  d.C=C;
  d.S=S;
  d.nsub=s.size();
  d.sub=s.data();
  gpu::check(fftwpp_gpu_plan_create(&d,&gpuplan),"plan creation");
  fftwpp_gpu_plan_set_tag(gpuplan,gputag);
  return gpuplan;
}

void fftBase::setTag(int tag)
{
  gputag=tag;
  if(gpuplan) fftwpp_gpu_plan_set_tag(gpuplan,gputag);
  if(planA) fftwpp_gpu_plan_set_tag(planA,gputag);
  if(planB) fftwpp_gpu_plan_set_tag(planB,gputag);
}

bool fftBase::innerEligible(Kind kind, size_t L, size_t m, size_t p, size_t C,
                            size_t S)
{
  return kind == COMPLEX && C == 1 && S == 1 && p > 2 && ispow2(m) &&
    ispow2(p) && m >= 16 && m <= 8192 && p >= 16 && p <= 4096 && L == p*m;
}

// Stage A: the reference's "L'=p, M'=q, m'=p, p'=1, q'=n" transform
// (convolve.cc:619) along t over the m columns s of f[t*m+s], followed by the
// outer twiddle zeta_N^{(n*u+r)*s}.
fftwpp_gpu_plan *fftBase::innerA()
{
  if(planA) return planA;
  std::vector<fftwpp_gpu_subblock> sub(n);
  for(size_t r=0; r < n; ++r) {
    memset(&sub[r],0,sizeof(sub[r]));
    sub[r].mlen=(uint32_t) p;
    sub[r].nout=(uint32_t) p;
    sub[r].k0=r;
    sub[r].off_call=sub[r].off_all=r*p*m;
  }
  fftwpp_gpu_pad_desc d;
  memset(&d,0,sizeof(d));
  d.kind=FFTWPP_KIND_COMPLEX;
  d.L=p;
  d.Lin=p;
  d.N=q;
  d.m=p;
  d.C=m;
  d.S=m;
  d.nsub=sub.size();
  d.sub=sub.data();
  gpu::check(fftwpp_gpu_plan_create(&d,&planA),"plan creation (inner A)");
  gpu::check(fftwpp_gpu_plan_set_outer(planA,plan(),n),"outer twiddle");
  fftwpp_gpu_plan_set_tag(planA,gputag);
  return planA;
}

// Stage B: q independent explicit (unpadded) length-m rows.
fftwpp_gpu_plan *fftBase::innerB()
{
  if(planB) return planB;
  fftwpp_gpu_subblock sb;
  memset(&sb,0,sizeof(sb));
  sb.mlen=(uint32_t) m;
  sb.nout=(uint32_t) m;
  fftwpp_gpu_pad_desc d;
  memset(&d,0,sizeof(d));
  d.kind=FFTWPP_KIND_COMPLEX;
  d.L=m;
  d.Lin=m;
  d.N=m;
  d.m=m;
  d.C=1;
  d.S=1;
  d.nsub=1;
  d.sub=&sb;
  gpu::check(fftwpp_gpu_plan_create(&d,&planB),"plan creation (inner B)");
  fftwpp_gpu_plan_set_tag(planB,gputag);
  return planB;
}

// Host-pointer staging for forward()/backward().
void fftBase::forward(Complex *f, Complex *F, size_t r, Complex *)
{
  const ResidueCall& c=call(r);
  size_t inBytes=doubles()*sizeof(double);
  size_t outBytes=(kind() == HERMITIAN ? 2*b*D*sizeof(double)
                   : outputSize()*sizeof(Complex));
  void *st=gpu::stream();
  bool fdev=gpu::isDevice(f), Fdev=gpu::isDevice(F);
  void *df=f, *dF=F;
  if(!fdev) {
    if(!devIn) gpu::check(fftwpp_gpu_malloc(&devIn,inBytes),"device allocation");
    gpu::check(fftwpp_gpu_memcpy_h2d(devIn,f,inBytes,st),"h2d");
    df=devIn;
  }
  if(!Fdev) {
    if(!devOut) gpu::check(fftwpp_gpu_malloc(&devOut,outBytes),"device allocation");
    dF=devOut;
  }
  gpu::check(fftwpp_gpu_forward(plan(),c.sb0,c.nsb,0,df,dF,1,0,0,st),
             "forward");
  if(!Fdev) {
    gpu::check(fftwpp_gpu_memcpy_d2h(F,devOut,outBytes,st),"d2h");
    gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  }
}

void fftBase::backward(Complex *F, Complex *f, size_t r, Complex *)
{
  const ResidueCall& c=call(r);
  size_t inBytes=doubles()*sizeof(double);
  size_t outBytes=(kind() == HERMITIAN ? 2*b*D*sizeof(double)
                   : outputSize()*sizeof(Complex));
  void *st=gpu::stream();
  bool fdev=gpu::isDevice(f), Fdev=gpu::isDevice(F);
  void *df=f, *dF=F;
  int accumulate=r > 0; // first block assigns, later blocks add
                        // (reference convolve.cc:1498-1502)
  if(!fdev) {
    if(!devIn) gpu::check(fftwpp_gpu_malloc(&devIn,inBytes),"device allocation");
    if(accumulate)
      gpu::check(fftwpp_gpu_memcpy_h2d(devIn,f,inBytes,st),"h2d");
    df=devIn;
  }
  if(!Fdev) {
    if(!devOut) gpu::check(fftwpp_gpu_malloc(&devOut,outBytes),"device allocation");
    gpu::check(fftwpp_gpu_memcpy_h2d(devOut,F,outBytes,st),"h2d");
    dF=devOut;
  }
  gpu::check(fftwpp_gpu_backward(plan(),c.sb0,c.nsb,0,dF,df,accumulate,1.0,
                                 1,0,0,st),"backward");
  if(!fdev) {
    if(S == C)
      gpu::check(fftwpp_gpu_memcpy_d2h(f,devIn,inBytes,st),"d2h");
    else // preserve the caller's stride gaps
      gpu::check(fftwpp_gpu_memcpy2d(f,S*wordSize()*sizeof(double),devIn,
                                     S*wordSize()*sizeof(double),
                                     C*wordSize()*sizeof(double),
                                     inputLength(),1,st),"d2h");
    gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  }
}

double fftBase::time()
{
  Convolution conv(this);
  size_t N=std::max(app.A,app.B);
  DeviceArrays d;
  d.ensure(N,doubles()*sizeof(double));
  std::vector<Complex *> f(N);
  for(size_t a=0; a < N; ++a) {
    gpu::check(fftwpp_gpu_memset(d.ptr[a],0,doubles()*sizeof(double),
                                 gpu::stream()),"memset");
    f[a]=(Complex *) d.ptr[a];
  }
  std::vector<double> T;
  for(int it=0; it < 7; ++it) {
    gpu::check(fftwpp_gpu_stream_sync(gpu::stream()),"sync");
    auto t0=std::chrono::steady_clock::now();
    conv.convolveRaw(f.data());
    gpu::check(fftwpp_gpu_stream_sync(gpu::stream()),"sync");
    auto t1=std::chrono::steady_clock::now();
    T.push_back(std::chrono::duration<double,std::nano>(t1-t0).count());
  }
  std::sort(T.begin(),T.end());
  return T[T.size()/2];
}

double fftBase::report()
{
  double median=time()*1.0e-9;
  std::cout << "median=" << median << std::endl;
  return median;
}

// ---------------------------------------------------------------------------
// fftPad / fftPadCentered
// ---------------------------------------------------------------------------

fftPad::fftPad(size_t L, size_t M, Application& app, size_t C, size_t S,
               bool Explicit) :
  fftBase(L,M,app,C,S,false)
{
  choose(Explicit);
  if(Explicit) this->M=m;
  init();
}

fftPad::fftPad(size_t L, size_t M, Application &app, size_t C, size_t S,
               size_t m, size_t D, bool inplace) :
  fftBase(L,M,app,C,S,m,D,inplace,false)
{
  parameters(L,M,m,centered,p,n,q);
  if(q > 1 && !valid(m,p,q,n,D,this->S)) invalid();
  init();
}

fftPad::fftPad(size_t L, size_t M, Application &app, size_t C, size_t S,
               Deferred) :
  fftBase(L,M,app,C,S,true) {}

fftPad::fftPad(size_t L, size_t M, Application &app, size_t C, size_t S,
               size_t m, size_t D, bool inplace, Deferred) :
  fftBase(L,M,app,C,S,m,D,inplace,true) {}

fftPadCentered::fftPadCentered(size_t L, size_t M, Application& app, size_t C,
                               size_t S, bool Explicit) :
  fftPad(L,M,app,C,S,Deferred())
{
  choose(Explicit);
  if(Explicit) this->M=m;
  init();
}

fftPadCentered::fftPadCentered(size_t L, size_t M, Application &app, size_t C,
                               size_t S, size_t m, size_t D, bool inplace) :
  fftPad(L,M,app,C,S,m,D,inplace,Deferred())
{
  parameters(L,M,m,centered,p,n,q);
  if(q > 1 && !valid(m,p,q,n,D,this->S)) invalid();
  init();
}

void fftPad::init()
{
  common();
  std::vector<SubBlockHost> sub;
  callTable.clear();
  totalRows=0;
  if(q == 1) {
    dr=D=D0=R=1;
    l=M;
    b=S*l;
    SubBlockHost h;
    memset(&h,0,sizeof(h));
    h.s.mlen=(uint32_t) m;
    h.s.nout=(uint32_t) m;
    h.s.k0=0;
    sub.push_back(h);
    ResidueCall c={0,0,1,m,0};
    callTable.push_back(c);
    totalRows=m;
  } else {
    size_t P=(p == 2) ? 1 : (centered ? p/2 : p);
    if(centered && p % 2 != 0) {
      std::cerr << "Odd values of p are incompatible with the centered and "
                << "Hermitian routines." << std::endl;
      invalid();
    }
    l=m*P;
    b=S*l;
    dr=Dr();
    R=residueBlocks();
    D0=n % D;
    if(D0 == 0) D0=D;
    const size_t N=m*q;
    for(size_t r=0; r < R; r += increment(r)) {
      size_t blocks=(r == 0) ? D0 : D;
      ResidueCall c={r,sub.size(),0,0,totalRows};
      for(size_t d=0; d < blocks; ++d) {
        for(size_t u=0; u < P; ++u) {
          size_t i0=d*l+u*m;
          size_t k0=index(r,i0);
          // consecutive positions advance the global index by q
          if(m > 1 && index(r,i0+1) != (k0+q) % N) {
            std::cerr << "internal error: residue layout mismatch at r=" << r
                      << " i=" << i0 << std::endl;
            exit(-1);
          }
          SubBlockHost h;
          memset(&h,0,sizeof(h));
          h.s.mlen=(uint32_t) m;
          h.s.nout=(uint32_t) m;
          h.s.k0=k0;
          h.s.off_call=b*d+S*m*u;
          h.s.off_all=S*totalRows;
          sub.push_back(h);
          totalRows += m;
          c.rows += m;
          ++c.nsb;
        }
      }
      callTable.push_back(c);
    }
  }
  buildPlan(sub);
  twoStage=q > 1 && innerEligible(kind(),L,m,p,C,S) &&
    (m <= 4096 || longStageB(app));
  report(centered ? "fftPadCentered" : "fftPad");
}

// ---------------------------------------------------------------------------
// fftPadHermitian
// ---------------------------------------------------------------------------

fftPadHermitian::fftPadHermitian(size_t L, size_t M, Application& app,
                                 size_t C, bool Explicit) :
  fftBase(L,M,app,C,C,true)
{
  choose(Explicit);
  if(Explicit) this->M=m;
  init();
}

fftPadHermitian::fftPadHermitian(size_t L, size_t M, Application &app,
                                 size_t C, size_t m, size_t D, bool inplace) :
  fftBase(L,M,app,C,C,m,D,inplace,true)
{
  parameters(L,M,m,centered,p,n,q);
  if(q > 1 && !valid(m,p,q,n,D,C)) invalid();
  init();
}

void fftPadHermitian::init()
{
  common();
  S=C; // stride gaps are not supported for Hermitian transforms
  e=m/2+1;
  std::vector<SubBlockHost> sub;
  callTable.clear();
  totalRows=0;
  if(q == 1) {
    B=b=C*e;
    dr=D0=R=1;
    D=1;
    l=m;
    SubBlockHost h;
    memset(&h,0,sizeof(h));
    h.s.mlen=(uint32_t) m;
    h.s.nout=(uint32_t) m;
    sub.push_back(h);
    ResidueCall c={0,0,1,m,0};
    callTable.push_back(c);
    totalRows=m;
  } else {
    if(p % 2 != 0) {
      std::cerr << "Odd values of p are incompatible with the centered and "
                << "Hermitian routines." << std::endl;
      invalid();
    }
    dr=Dr();
    size_t p2=p/2;
    b=align(ceilquotient(p2*Cm,2));
    B=align(p2*C*e);
    if(inplace) b=B;
    l=m*p2;
    R=residueBlocks();
    D0=n % D;
    if(D0 == 0) D0=D;
    const size_t stride=blocksize(0);
    const size_t N=m*q;
    for(size_t r=0; r < R; r += increment(r)) {
      size_t blocks=(r == 0) ? D0 : D;
      ResidueCall c={r,sub.size(),0,0,totalRows};
      for(size_t d=0; d < blocks; ++d) {
        // p > 2 (reference forwardInner, convolve.cc:5014-5230): p/2 real
        // blocks of m outputs per residue, one per inner index u
        for(size_t u=0; u < p2; ++u) {
          size_t i0=stride*d+u*m;
          SubBlockHost h;
          memset(&h,0,sizeof(h));
          h.s.mlen=(uint32_t) m;
          h.s.nout=(uint32_t) m;
          h.s.k0=index(r,i0);
          if(m > 1 && index(r,i0+1) != (h.s.k0+q) % N) {
            std::cerr << "internal error: Hermitian residue layout mismatch at"
                      << " r=" << r << " i=" << i0 << std::endl;
            exit(-1);
          }
          h.s.off_call=2*b*d+C*m*u; // doubles
          h.s.off_all=C*totalRows;  // doubles
          sub.push_back(h);
          totalRows += m;
          c.rows += m;
          ++c.nsb;
        }
      }
      callTable.push_back(c);
    }
  }
  buildPlan(sub);
  report("fftPadHermitian");
}

// ---------------------------------------------------------------------------
// fftPadReal
// ---------------------------------------------------------------------------

fftPadReal::fftPadReal(size_t L, size_t M, Application& app, size_t C,
                       size_t S, bool Explicit) :
  fftBase(L,M,app,C,S,false)
{
  choose(Explicit);
  if(Explicit) this->M=m;
  init();
}

fftPadReal::fftPadReal(size_t L, size_t M, Application &app, size_t C,
                       size_t S, size_t m, size_t D, bool inplace) :
  fftBase(L,M,app,C,S,m,D,inplace,false)
{
  parameters(L,M,m,centered,p,n,q);
  if(q > 1 && !valid(m,p,q,n,D,this->S)) invalid();
  init();
}

// Restates reference convolve.h:956-979 (sign -1 / r2c index convention).
size_t fftPadReal::index(size_t r, size_t i)
{
  if(q == 1) return i;
  const size_t N=q*m;
  size_t s=i % m;
  size_t P=p == 2 ? 1 : p;
  r += i/(P*m);
  if(p <= 2) {
    if(r == 0) return q*i;
    if(2*r == q) return N-(2*q*i+r);
  } else {
    size_t u=(i/m) % p;
    if(r == 0) {
      if(2*u == p) return N-(2*q*s+u*n);
      return s == 0 ? u*n : N-(q*s-u*n);
    }
    if(2*r == n) return N-(q*s+2*u*n+r);
    return q*(m-s)-(u*n+r);
  }
  return q*(m-s)-r;
}

void fftPadReal::init()
{
  common();
  e=m/2+1;
  std::vector<SubBlockHost> sub;
  callTable.clear();
  totalRows=0;
  if(q == 1) {
    l=e;
    b=S*l;
    dr=D0=R=n=1;
    D=1;
    SubBlockHost h;
    memset(&h,0,sizeof(h));
    h.s.mlen=(uint32_t) m;
    h.s.nout=(uint32_t) e;
    h.s.flags=FFTWPP_SB_CONJ_OUT;
    sub.push_back(h);
    ResidueCall c={0,0,1,e,0};
    callTable.push_back(c);
    totalRows=e;
  } else {
    const size_t P=p == 2 ? 1 : p;
    const size_t N=m*q;
    l=m*P;
    b=S*l;
    dr=Dr();
    R=residueBlocks();
    D0=((n-1)/2) % D;
    if(D0 == 0) D0=D;
    // Position i of a call holds the reference's value at index(r,i), which
    // for real data equals the sign +1 transform at k_i=(N-index(r,i)) mod N
    // (reference tests/hybridr.cc:88-90).  Apart from the r2c block of
    // p <= 2, every block is a natural-order sub-block: k_i=k0+(N/mlen)*s.
    auto add=[&](ResidueCall& c, size_t r, size_t i0, size_t mlen,
                 uint32_t flags) {
      SubBlockHost h;
      memset(&h,0,sizeof(h));
      h.s.mlen=(uint32_t) mlen;
      h.s.nout=(uint32_t) mlen;
      h.s.flags=flags;
      size_t k0=(N-index(r,i0)) % N;
      if(mlen > 1 && (N-index(r,i0+1)) % N != (k0+N/mlen) % N) {
        std::cerr << "internal error: real residue layout mismatch at r=" << r
                  << " i=" << i0 << std::endl;
        exit(-1);
      }
      h.s.k0=k0;
      h.s.off_call=S*i0;
      h.s.off_all=S*totalRows;
      sub.push_back(h);
      totalRows += mlen;
      c.rows += mlen;
      ++c.nsb;
    };
    for(size_t r=0; r < R; r += increment(r)) {
      ResidueCall c={r,sub.size(),0,0,totalRows};
      if(r == 0) {
        if(p <= 2) { // r2c: e outputs stored with the sign -1 convention
          SubBlockHost h;
          memset(&h,0,sizeof(h));
          h.s.mlen=(uint32_t) m;
          h.s.nout=(uint32_t) e;
          h.s.flags=FFTWPP_SB_CONJ_OUT;
          h.s.k0=0;
          h.s.off_call=0;
          h.s.off_all=S*totalRows;
          sub.push_back(h);
          totalRows += e;
          c.rows += e;
          ++c.nsb;
        } else {
          // residues -u*n, u <= p/2 (reference forwardInner,
          // convolve.cc:6168-6330): u=0 holds both halves of its spectrum;
          // for even p the class 2u == p is the packed half-length one
          for(size_t u=0; u < ceilquotient(p,2); ++u)
            add(c,0,u*m,m,u == 0 ? FFTWPP_SB_SELFCONJ : 0);
          if(p % 2 == 0) add(c,0,(p/2)*m,e-1,0);
        }
      } else if(2*r == n) {
        if(p <= 2) add(c,r,0,e-1,0);  // packed half-length class
        else
          for(size_t u=0; u < p/2; ++u) add(c,r,u*m,m,0);
      } else {
        size_t blocks=(r == 1) ? D0 : D;
        for(size_t d=0; d < blocks; ++d)
          for(size_t u=0; u < P; ++u)
            add(c,r,d*l+u*m,m,0);
      }
      callTable.push_back(c);
    }
  }
  buildPlan(sub);
  report("fftPadReal");
}

// ---------------------------------------------------------------------------
// Convolution (1-D)
// ---------------------------------------------------------------------------

Convolution::Convolution(fftBase *fft, Complex **, Complex *, Complex *) :
  ThreadBase(fft->Threads()), fft(fft), L(fft->L), A(fft->app.A),
  B(fft->app.B), mult(fft->app.mult)
{
  indices.copy(NULL,0);
  indices.fft=fft;
  rowIndexDims=0;
  scale=1.0/normalization();
  if(fft->tag() == 0) fft->setTag(1);
  multId=multiplierId(mult);
  if(std::max(A,B) > MAXFUSED && multId >= 0) {
    // the fused kernels take at most MAXFUSED arrays; larger A or B run the
    // unfused forward / multiply / backward path (the built-ins have host
    // bodies, so they can serve as "user" multipliers there)
    multId=-1;
  }
  if(fft->C != 1 && multId != FFTWPP_MULT_NONE) {
    // as in the reference the multiplier only sees C == 1 data
  }
}

Convolution::~Convolution() {}

void Convolution::convolveRows(Complex **f, size_t offset, size_t nrows,
                               size_t rowstride, double sc)
{
  if(multId < 0) {
    runCustom(f,offset,nrows,rowstride,sc);
    return;
  }
  size_t N=std::max(A,B);
  void *ptrs[MAXFUSED];
  if(fft->C != 1 && multId == FFTWPP_MULT_NONE) {
    // C interleaved columns without a multiplier (what fft->report()/time()
    // and tests/hybrid*.cc -C run): forward and backward passes over all
    // residues; the fused row kernels need C == 1
    const std::vector<ResidueCall>& calls=fft->calls();
    size_t nsub=calls.back().sb0+calls.back().nsb;
    bool herm=fft->kind() == fftBase::HERMITIAN;
    size_t words=fft->allSize();
    size_t wbytes=herm ? sizeof(double) : sizeof(Complex);
    devT.ensure(1,nrows*words*wbytes);
    size_t inWord=fft->wordSize() == 1 ? 1 : 1; // rowstride is in input words
    (void) inWord;
    void *st=gpu::stream();
    for(size_t b=0; b < B; ++b) {
      gpu::check(fftwpp_gpu_forward(fft->plan(),0,nsub,1,f[b]+offset,
                                    devT.ptr[0],nrows,rowstride,words,st),
                 "forward");
      gpu::check(fftwpp_gpu_backward(fft->plan(),0,nsub,1,devT.ptr[0],
                                     f[b]+offset,0,sc,nrows,words,rowstride,
                                     st),"backward");
    }
    return;
  }
  if(fft->innerFast() && multId != FFTWPP_MULT_NONE) {
    // two-stage large transform: f -> T (stage A), q rows of length m fused
    // in T (stage B), T -> f (stage A adjoint, normalisation folded in)
    size_t q=fft->q, m=fft->m;
    size_t words=q*m;
    devT.ensure(N,nrows*words*sizeof(Complex));
    void *st=gpu::stream();
    for(size_t a=0; a < A; ++a)
      gpu::check(fftwpp_gpu_forward(fft->innerA(),0,fft->n,1,f[a]+offset,
                                    devT.ptr[a],nrows,rowstride,words,st),
                 "forward (inner A)");
    for(size_t a=0; a < N; ++a) ptrs[a]=devT.ptr[a];
    gpu::check(fftwpp_gpu_convolve(fft->innerB(),ptrs,(uint32_t) A,
                                   (uint32_t) B,multId,1.0,nrows*q,m,st),
               "convolve (inner B)");
    for(size_t b=0; b < B; ++b)
      gpu::check(fftwpp_gpu_backward(fft->innerA(),0,fft->n,1,devT.ptr[b],
                                     f[b]+offset,0,sc,nrows,words,rowstride,
                                     st),"backward (inner A)");
    return;
  }
  for(size_t a=0; a < N; ++a)
    ptrs[a]=(void *) (f[a]+offset);
  gpu::check(fftwpp_gpu_convolve(fft->plan(),ptrs,(uint32_t) A,(uint32_t) B,
                                 multId,sc,nrows,rowstride,gpu::stream()),
             "convolve");
}

// User multipliers are not fused: GPU forward of every residue -> multiplier
// per residue block (reference operate(), convolve.h:1120-1135) -> GPU
// backward.  With a registered device implementation the transformed data
// never leaves the GPU and nothing synchronises; otherwise the blocks make a
// round trip through host memory for the host function.
void Convolution::runCustom(Complex **f, size_t offset, size_t nrows,
                            size_t rowstride, double sc)
{
  size_t N=std::max(A,B);
  bool herm=fft->kind() == fftBase::HERMITIAN;
  size_t words=fft->allSize();
  size_t wbytes=herm ? sizeof(double) : sizeof(Complex);
  deviceMultiplier *dmult=deviceMultiplierOf(mult);
  DeviceArrays F;
  F.ensure(N,words*wbytes);
  std::vector<Complex *> hostF(N,(Complex *) NULL);
  if(!dmult)
    for(size_t a=0; a < N; ++a)
      hostF[a]=(Complex *) malloc(words*wbytes);
  const std::vector<ResidueCall>& calls=fft->calls();
  size_t nsub=calls.back().sb0+calls.back().nsb;
  void *st=gpu::stream();
  for(size_t row=0; row < nrows; ++row) {
    size_t off=offset+row*rowstride/(fft->wordSize() == 1 ? 2 : 1);
    if(fft->wordSize() == 1 && (row*rowstride) % 2) {
      std::cerr << "custom multipliers need even row strides for real data"
                << std::endl;
      exit(-1);
    }
    for(size_t a=0; a < A; ++a) {
      gpu::check(fftwpp_gpu_forward(fft->plan(),0,nsub,1,f[a]+off,F.ptr[a],
                                    1,0,0,st),"forward");
      if(!dmult)
        gpu::check(fftwpp_gpu_memcpy_d2h(hostF[a],F.ptr[a],words*wbytes,st),
                   "d2h");
    }
    if(!dmult) gpu::check(fftwpp_gpu_stream_sync(st),"sync");
    std::vector<Complex *> G(N);
    for(size_t ic=0; ic < calls.size(); ++ic) {
      const ResidueCall& c=calls[ic];
      size_t bs=fft->blocksize(c.r);
      // one multiplier call per sub-block group of `bs` outputs
      size_t done=0;
      size_t d=0;
      while(done < c.rows) {
        size_t rowsHere=std::min(bs,c.rows-done);
        for(size_t a=0; a < N; ++a) {
          Complex *base=dmult ? (Complex *) F.ptr[a] : hostF[a];
          G[a]=herm ? (Complex *) ((double *) base+fft->C*(c.row0+done))
            : base+fft->S*(c.row0+done);
        }
        indices.r=c.r;
        indices.offset=d*fft->b;
        // outer transformed indices of this batched row (reference
        // convolve.h:1442,1759 set them per row before the inner call)
        if(rowIndexDims && (row+1)*rowIndexDims <= rowIndex.size())
          for(size_t q=0; q < rowIndexDims && q < indices.maxsize; ++q)
            indices.index[q]=rowIndex[row*rowIndexDims+q];
        if(dmult) (*dmult)(G.data(),rowsHere,&indices,st);
        else (*mult)(G.data(),rowsHere,&indices,threads);
        done += rowsHere;
        ++d;
      }
    }
    for(size_t bq=0; bq < B; ++bq) {
      if(!dmult)
        gpu::check(fftwpp_gpu_memcpy_h2d(F.ptr[bq],hostF[bq],words*wbytes,st),
                   "h2d");
      gpu::check(fftwpp_gpu_backward(fft->plan(),0,nsub,1,F.ptr[bq],
                                     f[bq]+off,0,sc,1,0,0,st),"backward");
    }
    if(!dmult) gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  }
  // F is released at scope exit: the stream must have drained it first
  if(dmult) gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  for(size_t a=0; a < N; ++a) free(hostF[a]);
}

void Convolution::run(Complex **f, size_t offset, double sc)
{
  size_t N=std::max(A,B);
  if(gpu::isDevice(f[0])) {
    convolveRows(f,offset,1,0,sc);
    return;
  }
  size_t bytes=fft->doubles()*sizeof(double);
  dev.ensure(N,bytes);
  void *st=gpu::stream();
  std::vector<Complex *> d(N);
  for(size_t a=0; a < N; ++a) d[a]=(Complex *) dev.ptr[a];
  for(size_t a=0; a < A; ++a)
    gpu::check(fftwpp_gpu_memcpy_h2d(dev.ptr[a],f[a]+offset,bytes,st),"h2d");
  convolveRows(d.data(),0,1,0,sc);
  for(size_t bq=0; bq < B; ++bq)
    gpu::check(fftwpp_gpu_memcpy_d2h(f[bq]+offset,dev.ptr[bq],bytes,st),"d2h");
  gpu::check(fftwpp_gpu_stream_sync(st),"sync");
}

void Convolution::convolveRaw(Complex **f) {run(f,0,1.0);}

void Convolution::convolveRaw(Complex **f, Indices *indices2)
{
  indices.copy(indices2,0);
  indices.fft=fft;
  run(f,0,1.0);
}

void Convolution::convolveRaw(Complex **f, size_t offset) {run(f,offset,1.0);}

void Convolution::convolveRaw(Complex **f, size_t offset, Indices *indices2)
{
  indices.copy(indices2,0);
  indices.fft=fft;
  run(f,offset,1.0);
}

void Convolution::convolve(Complex **f) {run(f,0,scale);}
void Convolution::convolve(Complex **f, size_t offset) {run(f,offset,scale);}

static void scaleBox(Complex *h, double scale, size_t n0, size_t n1,
                     size_t n2, size_t s0, size_t s1)
{
  double *x=(double *) h;
  if(gpu::isDevice(h)) {
    gpu::check(fftwpp_gpu_scale(x,scale,n0,n1,n2,s0,s1,gpu::stream()),
               "scale");
    return;
  }
  for(size_t i=0; i < n0; ++i)
    for(size_t j=0; j < n1; ++j) {
      double *row=x+i*s0+j*s1;
      for(size_t k=0; k < n2; ++k)
        row[k] *= scale;
    }
}

void Convolution::normalize(Complex **h, size_t offset)
{
  size_t wL=fft->wordSize()*fft->inputLength();
  for(size_t bq=0; bq < B; ++bq)
    scaleBox(h[bq]+offset,scale,1,1,wL,0,0);
}

// ---------------------------------------------------------------------------
// Hermitian symmetrization (host data), reference convolve.h:1168-1267
// ---------------------------------------------------------------------------

void HermitianSymmetrizeX(size_t Hx, size_t Hy, size_t x0, Complex *f,
                          size_t Sx, size_t)
{
  Complex *origin=f+x0*Sx;
  for(size_t i=1; i < Hx; ++i)
    *(origin-i*Sx)=conj(origin[i*Sx]);
  origin[0]=Complex(origin[0].real(),0.0);
  if(x0 == Hx) // even length: zero the unpaired Nyquist row
    for(size_t j=0; j < Hy; ++j)
      f[j]=0.0;
}

void HermitianSymmetrizeXY(size_t Hx, size_t Hy, size_t Hz, size_t x0,
                           size_t y0, Complex *f, size_t Sx, size_t Sy,
                           size_t)
{
  size_t origin=x0*Sx+y0*Sy;
  Complex *F=f+origin;
  for(size_t i=1; i < Hx; ++i)
    *(F-i*Sx)=conj(F[i*Sx]);
  F[0]=Complex(F[0].real(),0.0);

  for(ptrdiff_t i=-(ptrdiff_t) Hx+1; i < (ptrdiff_t) Hx; ++i)
    for(size_t j=1; j < Hy; ++j)
      f[origin-i*(ptrdiff_t) Sx-j*Sy]=conj(f[origin+i*(ptrdiff_t) Sx+j*Sy]);

  if(x0 == Hx) {
    size_t Ly=y0+Hy;
    for(size_t j=0; j < Ly; ++j)
      for(size_t k=0; k < Hz; ++k)
        f[Sy*j+k]=0.0;
  }
  if(y0 == Hy) {
    size_t Lx=x0+Hx;
    for(size_t i=0; i < Lx; ++i)
      for(size_t k=0; k < Hz; ++k)
        f[Sx*i+k]=0.0;
  }
}

// ---------------------------------------------------------------------------
// Convolution2
// ---------------------------------------------------------------------------


static size_t envSize(const char *name, size_t def)
{
  const char *s=getenv(name);
  if(!s || !*s) return def;
  return (size_t) strtoull(s,NULL,10);
}

Convolution2::Convolution2(fftBase *fftx, fftBase *ffty, Complex **,
                           Complex *, Complex *) :
  ThreadBase(fftx->Threads()), fftx(fftx), ffty(ffty), A(fftx->app.A),
  B(fftx->app.B), mult(fftx->app.mult), planeChunk(0), FxBytes(0)
{
  threads=1;
  convolvey=new Convolution*[1];
  convolvey[0]=new Convolution(ffty);
  fftx->setTag(1);
  ffty->setTag(2);
  Lx=fftx->L;
  Ly=fftx->C;
  Sx=fftx->S;
  scale=1.0/normalization();
  indices.copy(NULL,1);
  planeChunk=envSize("FFTWPP_PLANE_CHUNK",0);
  if(fftx->kind() == fftBase::HERMITIAN) {
    std::cerr << "fftPadHermitian can only be the innermost dimension"
              << std::endl;
    exit(-1);
  }
}

Convolution2::~Convolution2()
{
  delete convolvey[0];
  delete [] convolvey;
}

// x pass over all residues at once, batched inner convolutions over every
// transformed row, x backward pass with the normalisation folded in.
void Convolution2::convolvePlanes(Complex **F, size_t offset, size_t nplanes,
                                  size_t planestride, double sc)
{
  // Here "fftx" is the strided pass of this 2-D object (the y pass of a 3-D
  // convolution) and each "plane" is one x row of the caller.
  size_t N=std::max(A,B);
  size_t rows=fftx->allRows();
  size_t wordsPerPlane=rows*fftx->S;
  size_t chunk=planeChunk ? std::min(planeChunk,nplanes) : nplanes;
  devF.ensure(N,chunk*wordsPerPlane*sizeof(Complex));
  const std::vector<ResidueCall>& calls=fftx->calls();
  size_t nsub=calls.back().sb0+calls.back().nsb;
  void *st=gpu::stream();
  std::vector<Complex *> G(N);
  for(size_t a=0; a < N; ++a) G[a]=(Complex *) devF.ptr[a];
  bool real=fftx->wordSize() == 1;
  for(size_t i0=0; i0 < nplanes; i0 += chunk) {
    size_t np=std::min(chunk,nplanes-i0);
    for(size_t a=0; a < A; ++a) {
      const char *src=(const char *) (F[a]+offset)+
        i0*planestride*(real ? sizeof(double) : sizeof(Complex));
      gpu::check(fftwpp_gpu_forward(fftx->plan(),0,nsub,1,src,devF.ptr[a],np,
                                    planestride,wordsPerPlane,st),"forward");
    }
    if(convolvey[0]->customMultiplier()) {
      // per batched row: index[0] = transformed index of this object's
      // strided dimension, index[1] = the caller's (plane) index
      bool outer=planeIndex.size() >= i0+np;
      size_t dims=outer ? 2 : 1;
      std::vector<size_t> table(np*rows*dims);
      size_t base=indexBase();
      for(size_t ic=0; ic < calls.size(); ++ic) {
        const ResidueCall& c=calls[ic];
        for(size_t k=0; k < c.rows; ++k) {
          size_t idx=fftx->index(c.r,k+base);
          for(size_t pl=0; pl < np; ++pl) {
            size_t row=pl*rows+c.row0+k;
            table[row*dims]=idx;
            if(outer) table[row*dims+1]=planeIndex[i0+pl];
          }
        }
      }
      convolvey[0]->indices.copy(NULL,dims);
      convolvey[0]->setRowIndices(table,dims);
    }
    convolvey[0]->convolveRows(G.data(),0,np*rows,fftx->S,1.0);
    for(size_t bq=0; bq < B; ++bq) {
      if(bq < outBase.size() && outBase[bq]) {
        // offset is in words of F; planes are planestride words apart
        size_t plane0=(planestride ? offset/planestride : 0)+i0;
        int rc=FFTWPP_GPU_EUNSUPPORTED;
        if(bq < outDests.size() && !outDests[bq].empty())
          rc=fftwpp_gpu_backward_dests(fftx->plan(),0,nsub,devF.ptr[bq],
                                       outDests[bq].data(),
                                       (int) outDests[bq].size(),plane0,sc,np,
                                       wordsPerPlane,st);
        if(rc == FFTWPP_GPU_EUNSUPPORTED)
          rc=fftwpp_gpu_backward_mapped(fftx->plan(),0,nsub,devF.ptr[bq],
                                        outBase[bq],outStride[bq],plane0,sc,
                                        np,wordsPerPlane,st);
        gpu::check(rc,"backward (fused exchange)");
        continue;
      }
      char *dst=(char *) (F[bq]+offset)+
        i0*planestride*(real ? sizeof(double) : sizeof(Complex));
      gpu::check(fftwpp_gpu_backward(fftx->plan(),0,nsub,1,devF.ptr[bq],dst,
                                     0,sc,np,wordsPerPlane,planestride,st),
                 "backward");
    }
  }
}

void Convolution2::run(Complex **f, size_t offset, double sc)
{
  size_t N=std::max(A,B);
  if(gpu::isDevice(f[0])) {
    convolvePlanes(f,offset,1,0,sc);
    return;
  }
  size_t bytes=fftx->wordSize()*sizeof(double)*Lx*Sx;
  dev.ensure(N,bytes);
  void *st=gpu::stream();
  std::vector<Complex *> d(N);
  for(size_t a=0; a < N; ++a) d[a]=(Complex *) dev.ptr[a];
  for(size_t a=0; a < A; ++a)
    gpu::check(fftwpp_gpu_memcpy_h2d(dev.ptr[a],f[a]+offset,bytes,st),"h2d");
  convolvePlanes(d.data(),0,1,0,sc);
  for(size_t bq=0; bq < B; ++bq)
    gpu::check(fftwpp_gpu_memcpy_d2h(f[bq]+offset,dev.ptr[bq],bytes,st),"d2h");
  gpu::check(fftwpp_gpu_stream_sync(st),"sync");
}

void Convolution2::convolveRaw(Complex **f, size_t offset, Indices *ind)
{
  convolvey[0]->indices.copy(ind,1);
  run(f,offset,1.0);
}

void Convolution2::convolve(Complex **f, size_t offset)
{
  run(f,offset,scale);
}

void Convolution2::normalize(Complex **h, size_t offset)
{
  size_t w=fftx->wordSize();
  for(size_t bq=0; bq < B; ++bq)
    scaleBox(h[bq]+offset,scale,1,Lx,w*inputLengthy(),0,w*Sx);
}

// ---------------------------------------------------------------------------
// Convolution3
// ---------------------------------------------------------------------------

Convolution3::Convolution3(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                           Complex **, Complex *, Complex *, bool mpi) :
  ThreadBase(fftx->Threads()), fftx(fftx), ffty(ffty), fftz(fftz),
  A(fftx->app.A), B(fftx->app.B), mult(fftx->app.mult)
{
  threads=1;
  convolvez=NULL;
  convolveyz=new Convolution2*[1];
  convolveyz[0]=mpi ? NULL : new Convolution2(ffty,fftz);
  fftx->setTag(1);
  ffty->setTag(2);
  fftz->setTag(3);
  Lx=fftx->L;
  Ly=ffty->L;
  Lz=ffty->C;
  Sx=fftx->S;
  Sy=ffty->S;
  scale=1.0;
  if(!mpi) {
    scale=1.0/(fftx->normalization()*ffty->normalization()*
               fftz->normalization());
    checkStrides();
    // keep the y/z intermediates of a batch of x rows inside the L2
    size_t def=envSize("FFTWPP_PLANE_CHUNK",0);
    convolveyz[0]->planeChunk=def;
  }
  indices.copy(NULL,2);
}

Convolution3::~Convolution3()
{
  if(convolveyz[0]) delete convolveyz[0];
  delete [] convolveyz;
}

void Convolution3::checkStrides()
{
  if(Sx < Ly*Sy) {
    std::cerr << "Sx cannot be less than Ly*Sy" << std::endl;
    exit(-1);
  }
  if(fftx->C != (contiguous() ? Ly*Sy : Lz)) {
    std::cerr << "fftx->C is invalid" << std::endl;
    exit(-1);
  }
}

void Convolution3::run(Complex **f, size_t offset, double sc)
{
  size_t N=std::max(A,B);
  void *st=gpu::stream();
  bool real=fftx->wordSize() == 1;
  size_t wordBytes=real ? sizeof(double) : sizeof(Complex);
  std::vector<Complex *> d(N);
  bool host=!gpu::isDevice(f[0]);
  size_t bytes=wordBytes*Lx*Sx;
  if(host) {
    dev.ensure(N,bytes);
    for(size_t a=0; a < N; ++a) d[a]=(Complex *) dev.ptr[a];
    for(size_t a=0; a < A; ++a)
      gpu::check(fftwpp_gpu_memcpy_h2d(dev.ptr[a],f[a]+offset,bytes,st),"h2d");
  } else
    for(size_t a=0; a < N; ++a) d[a]=f[a]+offset;

  size_t rows=fftx->allRows();
  devF.ensure(N,rows*Sx*sizeof(Complex));
  const std::vector<ResidueCall>& calls=fftx->calls();
  size_t nsub=calls.back().sb0+calls.back().nsb;
  std::vector<Complex *> G(N);
  for(size_t a=0; a < N; ++a) G[a]=(Complex *) devF.ptr[a];

  // x forward, all residues.  Non-contiguous y stride: one launch row per y.
  size_t nr=contiguous() ? 1 : Ly;
  size_t rs=contiguous() ? 0 : Sy;
  for(size_t a=0; a < A; ++a)
    gpu::check(fftwpp_gpu_forward(fftx->plan(),0,nsub,1,d[a],devF.ptr[a],nr,
                                  rs,rs,st),"forward");
  // every transformed x row is an independent y-z convolution
  if(convolveyz[0]->convolvey[0]->customMultiplier()) {
    std::vector<size_t>& pi=convolveyz[0]->planeIndex;
    pi.assign(rows,0);
    size_t base=indexBase();
    for(size_t ic=0; ic < calls.size(); ++ic)
      for(size_t k=0; k < calls[ic].rows; ++k)
        pi[calls[ic].row0+k]=fftx->index(calls[ic].r,k+base);
  }
  convolveyz[0]->convolvePlanes(G.data(),0,rows,Sx,1.0);
  for(size_t bq=0; bq < B; ++bq)
    gpu::check(fftwpp_gpu_backward(fftx->plan(),0,nsub,1,devF.ptr[bq],d[bq],
                                   0,sc,nr,rs,rs,st),"backward");
  if(host) {
    for(size_t bq=0; bq < B; ++bq)
      gpu::check(fftwpp_gpu_memcpy_d2h(f[bq]+offset,dev.ptr[bq],bytes,st),
                 "d2h");
    gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  }
}

void Convolution3::convolveRaw(Complex **f, size_t offset, Indices *ind)
{
  convolveyz[0]->indices.copy(ind,2);
  run(f,offset,1.0);
}

void Convolution3::convolve(Complex **f, size_t offset)
{
  run(f,offset,scale);
}

void Convolution3::normalize(Complex **h, size_t offset)
{
  size_t w=fftx->wordSize();
  for(size_t bq=0; bq < B; ++bq)
    scaleBox(h[bq]+offset,scale,Lx,inputLengthy(),w*inputLengthz(),w*Sx,w*Sy);
}

} // namespace fftwpp
