// refapi.cc -- TEST INFRASTRUCTURE: a C-callable shim over the UNMODIFIED
// reference classes (compiled from /root/reference by oracle/Makefile against
// oracle/fftw3_shim).  It lets the Python tests and bench.py (--impl reference
// and the cpu_baseline leg ONLY) drive the reference's own fftPad* /
// Convolution{,2,3} code (convolve.h:471-1893) as the checker and as the CPU
// baseline.  It contains no reference code; it only calls the public API the
// way tests/hybridconv*.cc do (e.g. tests/hybridconvr3.cc:46-52).

#include "convolve.h"
#include "direct.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

using namespace fftwpp;
using namespace utils;

namespace {

multiplier *pickMult(int id)
{
  switch(id) {
    case 0: return multNone;
    case 1: return multBinary;
    case 2: return realMultBinary;
    case 3: return multcorrelation;
  }
  return multNone;
}

struct Pad {
  Application *app;
  fftBase *fft;
  int kind;
};

fftBase *makePad(int kind, size_t L, size_t M, Application &app, size_t C,
                 size_t S, size_t m, size_t D, long I)
{
  bool forced=m > 0;
  switch(kind) {
    case 0:
      return forced ? new fftPad(L,M,app,C,S,m,D,I) : new fftPad(L,M,app,C,S);
    case 1:
      return forced ? (fftBase *) new fftPadCentered(L,M,app,C,S,m,D,I) :
        (fftBase *) new fftPadCentered(L,M,app,C,S);
    case 2:
      return forced ? (fftBase *) new fftPadHermitian(L,M,app,C,m,D,I) :
        (fftBase *) new fftPadHermitian(L,M,app,C);
    case 3:
      return forced ? (fftBase *) new fftPadReal(L,M,app,C,S,m,D,I) :
        (fftBase *) new fftPadReal(L,M,app,C,S);
  }
  return NULL;
}

} // namespace

extern "C" {

void ref_set_maxthreads(size_t threads)
{
  fftw::maxthreads=threads;
}

size_t ref_get_max_threads()
{
  return parallel::get_max_threads();
}

// kind: 0 fftPad, 1 fftPadCentered, 2 fftPadHermitian, 3 fftPadReal.
// m == 0 lets the reference optimizer choose (it times candidates).
void *ref_pad_create(int kind, size_t L, size_t M, size_t C, size_t S,
                     size_t m, size_t D, long I, size_t A, size_t B,
                     int mult, size_t threads)
{
  Pad *P=new Pad;
  P->kind=kind;
  P->app=new Application(A,B,pickMult(mult),threads,false,m,D,m > 0 ? I : -1);
  P->fft=makePad(kind,L,M,*P->app,C,S,m,D,I);
  return P;
}

void ref_pad_destroy(void *h)
{
  Pad *P=(Pad *) h;
  delete P->fft;
  delete P->app;
  delete P;
}

// out[0..31]
void ref_pad_info(void *h, size_t *out)
{
  fftBase *f=((Pad *) h)->fft;
  size_t i=0;
  out[i++]=f->L; out[i++]=f->M; out[i++]=f->C; out[i++]=f->S;
  out[i++]=f->m; out[i++]=f->p; out[i++]=f->q; out[i++]=f->n;
  out[i++]=f->R; out[i++]=f->dr; out[i++]=f->D; out[i++]=f->D0;
  out[i++]=f->l; out[i++]=f->b; out[i++]=f->inplace; out[i++]=f->overwrite;
  out[i++]=f->centered; out[i++]=f->inputLength(); out[i++]=f->wordSize();
  out[i++]=f->doubles(); out[i++]=f->outputSize(); out[i++]=f->workSizeW();
  out[i++]=f->workSizeV(); out[i++]=f->nloops(); out[i++]=f->loop2();
  out[i++]=f->conjugates(); out[i++]=f->residueBlocks();
  out[i++]=f->paddedSize(); out[i++]=f->normalization(); out[i++]=f->repad();
  while(i < 32) out[i++]=0;
}

size_t ref_pad_increment(void *h, size_t r) {return ((Pad *) h)->fft->increment(r);}
size_t ref_pad_blocksize(void *h, size_t r) {return ((Pad *) h)->fft->blocksize(r);}
size_t ref_pad_noutputs(void *h, size_t r) {return ((Pad *) h)->fft->noutputs(r);}
size_t ref_pad_span(void *h, size_t r) {return ((Pad *) h)->fft->span(r);}
size_t ref_pad_index(void *h, size_t r, size_t i) {return ((Pad *) h)->fft->index(r,i);}

// Forward transform of residue block r: f (input words) -> F (outputSize()
// Complex words, caller allocated).  The input is copied first because some
// reference routines overwrite it.
void ref_pad_forward(void *h, const double *f, double *F, size_t r)
{
  fftBase *fft=((Pad *) h)->fft;
  size_t nd=fft->doubles();
  double *g=doubleAlign(nd);
  memcpy(g,f,nd*sizeof(double));
  size_t nW=fft->workSizeW();
  Complex *W=nW ? ComplexAlign(nW) : NULL;
  if(W) {
    memset((void *) W,0,nW*sizeof(Complex));
    fft->pad(W);
  }
  size_t nF=fft->outputSize();
  Complex *G=ComplexAlign(nF);
  memset((void *) G,0,nF*sizeof(Complex));
  fft->forward((Complex *) g,G,r,W);
  memcpy(F,G,nF*sizeof(Complex));
  if(fft->overwrite)
    memcpy((void *) f,g,nd*sizeof(double));
  deleteAlign(G);
  if(W) deleteAlign(W);
  deleteAlign(g);
}

// Backward transform of residue block r: F -> f (accumulating for r > 0 as the
// reference does).
void ref_pad_backward(void *h, const double *F, double *f, size_t r)
{
  fftBase *fft=((Pad *) h)->fft;
  size_t nd=fft->doubles();
  double *g=doubleAlign(nd);
  memcpy(g,f,nd*sizeof(double));
  size_t nW=fft->workSizeW();
  Complex *W=nW ? ComplexAlign(nW) : NULL;
  if(W) {
    memset((void *) W,0,nW*sizeof(Complex));
    fft->pad(W);
  }
  size_t nF=fft->outputSize();
  Complex *G=ComplexAlign(nF);
  memcpy((void *) G,F,nF*sizeof(Complex));
  fft->backward(G,(Complex *) g,r,W);
  memcpy(f,g,nd*sizeof(double));
  deleteAlign(G);
  if(W) deleteAlign(W);
  deleteAlign(g);
}

// A convolution bundle built exactly like tests/hybridconv{,h,r}{,2,3}.cc do.
struct Conv {
  int dim;
  Application *app[3];
  fftBase *fft[3];
  Convolution *c1;
  Convolution2 *c2;
  Convolution3 *c3;
  size_t A,B;
  size_t doubles; // doubles per input array
};

// family: 0 complex (fftPad...), 1 Hermitian (Centered...,Hermitian last),
//         2 real (fftPadReal first, fftPad after).
// L,M,m,D,I: per dimension (x,y,z order, dim entries); m[d]==0 => optimizer.
// Sx,Sy: strides (0 => contiguous).
void *ref_conv_create(int dim, int family, const size_t *L, const size_t *M,
                      const size_t *m, const size_t *D, const long *I,
                      size_t Sx, size_t Sy, size_t A, size_t B, int mult,
                      size_t threads, int verbose)
{
  Conv *c=new Conv;
  c->dim=dim;
  c->A=A; c->B=B;
  c->c1=NULL; c->c2=NULL; c->c3=NULL;
  for(int d=0; d < 3; ++d) {c->app[d]=NULL; c->fft[d]=NULL;}
  fftw::maxthreads=threads;

  int kinds[3];
  for(int d=0; d < dim; ++d) {
    if(family == 0) kinds[d]=0;
    else if(family == 1) kinds[d]=(d == dim-1) ? 2 : 1;
    else kinds[d]=(d == 0) ? 3 : 0;
  }

  size_t len[3]; // per-dimension stored length
  for(int d=0; d < dim; ++d)
    len[d]=(family == 1 && d == dim-1) ? ceilquotient(L[d],2) : L[d];

  for(int d=0; d < dim; ++d) {
    multiplier *mu=(d == dim-1) ? pickMult(mult) : multNone;
    long Id=m[d] > 0 ? I[d] : -1;
    if(d == 0)
      c->app[d]=new Application(A,B,mu,threads,verbose,m[d],D[d],Id);
    else
      c->app[d]=new Application(A,B,mu,*c->app[d-1],m[d],D[d],Id);
  }

  if(dim == 1) {
    c->fft[0]=makePad(kinds[0],L[0],M[0],*c->app[0],1,0,m[0],D[0],I[0]);
    c->c1=new Convolution(c->fft[0]);
    c->doubles=c->fft[0]->doubles();
  } else if(dim == 2) {
    if(Sx == 0) Sx=len[1];
    c->fft[0]=makePad(kinds[0],L[0],M[0],*c->app[0],len[1],Sx,m[0],D[0],I[0]);
    c->fft[1]=makePad(kinds[1],L[1],M[1],*c->app[1],1,0,m[1],D[1],I[1]);
    c->c2=new Convolution2(c->fft[0],c->fft[1]);
    c->doubles=c->fft[0]->wordSize()*L[0]*Sx;
  } else {
    if(Sy == 0) Sy=len[2];
    if(Sx == 0) Sx=L[1]*Sy;
    size_t Cx=(Sy == len[2] || kinds[0] == 3) ? L[1]*Sy : len[2];
    c->fft[0]=makePad(kinds[0],L[0],M[0],*c->app[0],Cx,Sx,m[0],D[0],I[0]);
    c->fft[1]=makePad(kinds[1],L[1],M[1],*c->app[1],len[2],Sy,m[1],D[1],I[1]);
    c->fft[2]=makePad(kinds[2],L[2],M[2],*c->app[2],1,0,m[2],D[2],I[2]);
    c->c3=new Convolution3(c->fft[0],c->fft[1],c->fft[2]);
    c->doubles=c->fft[0]->wordSize()*L[0]*Sx;
  }
  return c;
}

void ref_conv_destroy(void *h)
{
  Conv *c=(Conv *) h;
  delete c->c1; delete c->c2; delete c->c3;
  for(int d=2; d >= 0; --d) {
    delete c->fft[d];
    delete c->app[d];
  }
  delete c;
}

// Parameters actually used in dimension d: out = {m,p,q,n,D,inplace,C,S}
void ref_conv_params(void *h, int d, size_t *out)
{
  Conv *c=(Conv *) h;
  fftBase *f=c->fft[d];
  out[0]=f->m; out[1]=f->p; out[2]=f->q; out[3]=f->n; out[4]=f->D;
  out[5]=f->inplace; out[6]=f->C; out[7]=f->S;
}

size_t ref_conv_doubles(void *h) {return ((Conv *) h)->doubles;}

// Convolve A arrays (host pointers, each ref_conv_doubles() doubles) in place;
// the B outputs overwrite f[0..B).  normalized != 0 => convolve(), else
// convolveRaw() (tests/hybridconv.cc:83-96).
void ref_conv_convolve(void *h, double **f, int normalized)
{
  Conv *c=(Conv *) h;
  size_t N=std::max(c->A,c->B);
  double **g=doubleAlign(N,c->doubles);
  for(size_t a=0; a < c->A; ++a)
    memcpy(g[a],f[a],c->doubles*sizeof(double));
  Complex **G=(Complex **) g;
  if(c->c1) {if(normalized) c->c1->convolve(G); else c->c1->convolveRaw(G);}
  if(c->c2) {if(normalized) c->c2->convolve(G); else c->c2->convolveRaw(G);}
  if(c->c3) {if(normalized) c->c3->convolve(G); else c->c3->convolveRaw(G);}
  for(size_t b=0; b < c->B; ++b)
    memcpy(f[b],g[b],c->doubles*sizeof(double));
  deleteAlign(g[0]);
  delete [] g;
}

// Time `count` calls of convolveRaw on zero-filled data, as the reference's
// timing loops do (tests/hybridconv.cc:60,83-96); seconds[i] per call.
void ref_conv_time(void *h, size_t count, double *seconds)
{
  Conv *c=(Conv *) h;
  size_t N=std::max(c->A,c->B);
  double **g=doubleAlign(N,c->doubles);
  for(size_t a=0; a < N; ++a)
    memset(g[a],0,c->doubles*sizeof(double));
  Complex **G=(Complex **) g;
  for(size_t i=0; i < count; ++i) {
    auto t0=std::chrono::steady_clock::now();
    if(c->c1) c->c1->convolveRaw(G);
    if(c->c2) c->c2->convolveRaw(G);
    if(c->c3) c->c3->convolveRaw(G);
    auto t1=std::chrono::steady_clock::now();
    seconds[i]=std::chrono::duration<double>(t1-t0).count();
  }
  deleteAlign(g[0]);
  delete [] g;
}

// The reference's FFT-free direct convolutions (tests/direct.h, direct.cc).
void ref_direct_complex(int dim, const size_t *L, const double *f,
                        const double *g, double *h)
{
  Complex *F=(Complex *) f, *G=(Complex *) g, *H=(Complex *) h;
  if(dim == 1) {directconv<Complex> C(L[0]); C.convolve(H,F,G);}
  else if(dim == 2) {directconv2<Complex> C(L[0],L[1]); C.convolve(H,F,G);}
  else {directconv3<Complex> C(L[0],L[1],L[2]); C.convolve(H,F,G);}
}

void ref_direct_centered1(size_t L, const double *f, const double *g,
                          double *h)
{
  directconv<Complex> C(L);
  C.convolveC((Complex *) h,(Complex *) f,(Complex *) g);
}

void ref_direct_real(int dim, const size_t *L, const double *f,
                     const double *g, double *h)
{
  double *F=(double *) f, *G=(double *) g;
  if(dim == 1) {directconv<double> C(L[0]); C.convolve(h,F,G);}
  else if(dim == 2) {directconv2<double> C(L[0],L[1]); C.convolve(h,F,G);}
  else {directconv3<double> C(L[0],L[1],L[2]); C.convolve(h,F,G);}
}

// Hermitian direct convolutions; inputs must already be symmetrized.
void ref_direct_hermitian(int dim, const size_t *L, const double *f,
                          const double *g, double *h)
{
  Complex *F=(Complex *) f, *G=(Complex *) g, *H=(Complex *) h;
  if(dim == 1) {
    directconvh C(ceilquotient(L[0],2));
    C.convolve(H,F,G);
  } else if(dim == 2) {
    directconvh2 C(ceilquotient(L[0],2),ceilquotient(L[1],2),L[0]%2);
    C.convolve(H,F,G,false);
  } else {
    directconvh3 C(ceilquotient(L[0],2),ceilquotient(L[1],2),
                   ceilquotient(L[2],2),L[0]%2,L[1]%2);
    C.convolve(H,F,G,false);
  }
}

void ref_symmetrize(int dim, const size_t *L, double *f)
{
  Complex *F=(Complex *) f;
  if(dim == 1) HermitianSymmetrize(F);
  else if(dim == 2)
    HermitianSymmetrizeX(ceilquotient(L[0],2),ceilquotient(L[1],2),L[0]/2,F);
  else
    HermitianSymmetrizeXY(ceilquotient(L[0],2),ceilquotient(L[1],2),
                          ceilquotient(L[2],2),L[0]/2,L[1]/2,F);
}

}
