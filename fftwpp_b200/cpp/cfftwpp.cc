// cfftwpp.cc -- C-callable API (include/cfftwpp.h) over the host classes.
// Part 1 mirrors the entry points defined by reference wrappers/cfftw++.cc
// :27-163 and the owning bundles of wrappers/HybridConvolution.h:12-184
// (default M = A*L-A+1 for complex, 3*ceil(L/2)-2*(L%2) for Hermitian).
#include "convolve.h"
#include "mpiconvolve.h"
#include "mpifftw++.h"
#include "../../include/cfftwpp.h"
#include "../../include/fftwpp_gpu.h"

#include <algorithm>
#include <cstring>

using namespace fftwpp;
using namespace utils;

namespace {

multiplier *pickMult(int id)
{
  switch(id) {
    case 1: return multBinary;
    case 2: return realMultBinary;
    case 3: return multcorrelation;
  }
  return multNone;
}

fftBase *makePad(int kind, size_t L, size_t M, Application &app, size_t C,
                 size_t S, size_t m, size_t D, long I)
{
  // m alone (D == 0): the chooser honours Application.m and picks D itself,
  // as the reference optimizer does for a partially forced Application
  bool forced=m > 0 && D > 0;
  switch(kind) {
    case 0:
      return forced ? new fftPad(L,M,app,C,S,m,D,I != 0) :
        new fftPad(L,M,app,C,S);
    case 1:
      return forced ?
        (fftBase *) new fftPadCentered(L,M,app,C,S,m,D,I != 0) :
        (fftBase *) new fftPadCentered(L,M,app,C,S);
    case 2:
      return forced ? (fftBase *) new fftPadHermitian(L,M,app,C,m,D,I != 0) :
        (fftBase *) new fftPadHermitian(L,M,app,C);
    case 3:
      return forced ? (fftBase *) new fftPadReal(L,M,app,C,S,m,D,I != 0) :
        (fftBase *) new fftPadReal(L,M,app,C,S);
  }
  std::cerr << "unknown padded-FFT kind " << kind << std::endl;
  exit(-1);
}

struct Pad {
  Application *app;
  fftBase *fft;
};

// Generic convolution bundle (any dimension / family).
struct Conv {
  int dim;
  Application *app[3];
  fftBase *fft[3];
  Convolution *c1;
  Convolution2 *c2;
  Convolution3 *c3;
  size_t A,B;
  size_t doubles;

  // pipelined host-buffer entry: two slots of device staging, copy streams
  // and events (see fftwpp_conv_convolve_async)
  void *h2d,*d2h;
  void *evIn[2],*evDone[2],*evOut[2];
  bool busy[2];
  DeviceArrays slotBuf[2];

  Conv() : dim(0), c1(NULL), c2(NULL), c3(NULL), A(0), B(0), doubles(0),
           h2d(NULL), d2h(NULL) {
    for(int d=0; d < 3; ++d) {app[d]=NULL; fft[d]=NULL;}
    for(int s=0; s < 2; ++s) {
      evIn[s]=evDone[s]=evOut[s]=NULL;
      busy[s]=false;
    }
  }

  ~Conv() {
    for(int s=0; s < 2; ++s) {
      if(busy[s]) fftwpp_gpu_event_sync(evOut[s]);
      if(evIn[s]) fftwpp_gpu_event_destroy(evIn[s]);
      if(evDone[s]) fftwpp_gpu_event_destroy(evDone[s]);
      if(evOut[s]) fftwpp_gpu_event_destroy(evOut[s]);
    }
    if(h2d) fftwpp_gpu_stream_destroy(h2d);
    if(d2h) fftwpp_gpu_stream_destroy(d2h);
    delete c1; delete c2; delete c3;
    for(int d=2; d >= 0; --d) {
      delete fft[d];
      delete app[d];
    }
  }

  void convolve(Complex **f, bool normalized) {
    if(c1) {if(normalized) c1->convolve(f); else c1->convolveRaw(f);}
    if(c2) {if(normalized) c2->convolve(f); else c2->convolveRaw(f);}
    if(c3) {if(normalized) c3->convolve(f); else c3->convolveRaw(f);}
  }
};

Conv *makeConv(int dim, int family, const size_t *L, const size_t *M,
               const size_t *m, const size_t *D, const long *I, size_t Sx,
               size_t Sy, size_t A, size_t B, multiplier *mult)
{
  if(dim < 1 || dim > 3) {
    std::cerr << "dimension must be 1, 2 or 3" << std::endl;
    exit(-1);
  }
  Conv *c=new Conv;
  c->dim=dim;
  c->A=A;
  c->B=B;
  int kinds[3];
  size_t len[3];
  for(int d=0; d < dim; ++d) {
    if(family == 0) kinds[d]=0;
    else if(family == 1) kinds[d]=(d == dim-1) ? 2 : 1;
    else kinds[d]=(d == 0) ? 3 : 0;
    len[d]=(family == 1 && d == dim-1) ? ceilquotient(L[d],2) : L[d];
  }
  size_t zero[3]={0,0,0};
  long minus[3]={-1,-1,-1};
  if(!m) m=zero;
  if(!D) D=zero;
  if(!I) I=minus;
  for(int d=0; d < dim; ++d) {
    multiplier *mu=(d == dim-1) ? mult : multNone;
    long Id=m[d] > 0 ? I[d] : -1;
    if(d == 0)
      c->app[d]=new Application(A,B,mu,fftw::maxthreads,false,m[d],D[d],Id);
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

size_t defaultM(size_t L, size_t A) {return A*L-A+1;}
size_t defaultMh(size_t L) {return 3*ceilquotient(L,2)-2*(L % 2);}

Conv *simpleConv(int dim, int family, size_t Lx, size_t Ly, size_t Lz)
{
  size_t L[3]={Lx,Ly,Lz};
  size_t M[3];
  for(int d=0; d < dim; ++d)
    M[d]=family == 1 ? defaultMh(L[d]) : defaultM(L[d],2);
  return makeConv(dim,family,L,M,NULL,NULL,NULL,0,0,2,1,
                  pickMult(family == 1 ? 2 : 1));
}

void binary(Conv *c, fftwpp_cplx *a, fftwpp_cplx *b)
{
  Complex *F[]={(Complex *) a,(Complex *) b};
  c->convolve(F,true);
}

} // namespace

extern "C" {

double *create_doubleAlign(size_t n) {return doubleAlign(n);}
void delete_doubleAlign(double *p) {deleteAlign(p);}
fftwpp_cplx *create_complexAlign(size_t n)
{
  return (fftwpp_cplx *) ComplexAlign(n);
}
void delete_complexAlign(fftwpp_cplx *p) {deleteAlign(p);}

size_t get_fftwpp_maxthreads(void) {return fftw::maxthreads;}
void set_fftwpp_maxthreads(size_t nthreads) {fftw::maxthreads=nthreads;}

HybridConvolution *fftwpp_create_conv1d(size_t L)
{
  return (HybridConvolution *) simpleConv(1,0,L,0,0);
}
void fftwpp_conv1d_delete(HybridConvolution *conv) {delete (Conv *) conv;}
void fftwpp_conv1d_convolve(HybridConvolution *conv, fftwpp_cplx *a,
                            fftwpp_cplx *b)
{
  binary((Conv *) conv,a,b);
}

HybridConvolutionHermitian *fftwpp_create_hconv1d(size_t L)
{
  return (HybridConvolutionHermitian *) simpleConv(1,1,L,0,0);
}
void fftwpp_hconv1d_delete(HybridConvolutionHermitian *conv)
{
  delete (Conv *) conv;
}
void fftwpp_HermitianSymmetrize(fftwpp_cplx *f)
{
  HermitianSymmetrize((Complex *) f);
}
void fftwpp_hconv1d_convolve(HybridConvolutionHermitian *conv, fftwpp_cplx *a,
                             fftwpp_cplx *b)
{
  binary((Conv *) conv,a,b);
}

HybridConvolution2 *fftwpp_create_conv2d(size_t Lx, size_t Ly)
{
  return (HybridConvolution2 *) simpleConv(2,0,Lx,Ly,0);
}
void fftwpp_conv2d_delete(HybridConvolution2 *conv) {delete (Conv *) conv;}
void fftwpp_HermitianSymmetrizeX(size_t Hx, size_t Hy, size_t x0,
                                 fftwpp_cplx *f)
{
  HermitianSymmetrizeX(Hx,Hy,x0,(Complex *) f);
}
void fftwpp_conv2d_convolve(HybridConvolution2 *conv, fftwpp_cplx *a,
                            fftwpp_cplx *b)
{
  binary((Conv *) conv,a,b);
}
HybridConvolutionHermitian2 *fftwpp_create_hconv2d(size_t Lx, size_t Ly)
{
  return (HybridConvolutionHermitian2 *) simpleConv(2,1,Lx,Ly,0);
}
void fftwpp_hconv2d_delete(HybridConvolutionHermitian2 *conv)
{
  delete (Conv *) conv;
}
void fftwpp_hconv2d_convolve(HybridConvolutionHermitian2 *conv, fftwpp_cplx *a,
                             fftwpp_cplx *b)
{
  binary((Conv *) conv,a,b);
}

HybridConvolution3 *fftwpp_create_conv3d(size_t Lx, size_t Ly, size_t Lz)
{
  return (HybridConvolution3 *) simpleConv(3,0,Lx,Ly,Lz);
}
void fftwpp_conv3d_delete(HybridConvolution3 *conv) {delete (Conv *) conv;}
void fftwpp_HermitianSymmetrizeXY(size_t Hx, size_t Hy, size_t Hz, size_t x0,
                                  size_t y0, fftwpp_cplx *f)
{
  HermitianSymmetrizeXY(Hx,Hy,Hz,x0,y0,(Complex *) f);
}
void fftwpp_conv3d_convolve(HybridConvolution3 *conv, fftwpp_cplx *a,
                            fftwpp_cplx *b)
{
  binary((Conv *) conv,a,b);
}
HybridConvolutionHermitian3 *fftwpp_create_hconv3d(size_t Lx, size_t Ly,
                                                   size_t Lz)
{
  return (HybridConvolutionHermitian3 *) simpleConv(3,1,Lx,Ly,Lz);
}
void fftwpp_hconv3d_delete(HybridConvolutionHermitian3 *conv)
{
  delete (Conv *) conv;
}
void fftwpp_hconv3d_convolve(HybridConvolutionHermitian3 *conv, fftwpp_cplx *a,
                             fftwpp_cplx *b)
{
  binary((Conv *) conv,a,b);
}

// ---------------- generic handle API ----------------

void *fftwpp_pad_create(int kind, size_t L, size_t M, size_t C, size_t S,
                        size_t m, size_t D, long I, size_t A, size_t B,
                        int mult)
{
  Pad *P=new Pad;
  P->app=new Application(A,B,pickMult(mult),fftw::maxthreads,false,m,D,
                         m > 0 ? I : -1);
  P->fft=makePad(kind,L,M,*P->app,C,S,m,D,I);
  return P;
}

void fftwpp_pad_destroy(void *pad)
{
  Pad *P=(Pad *) pad;
  if(!P) return;
  delete P->fft;
  delete P->app;
  delete P;
}

void fftwpp_pad_info(void *pad, size_t *out)
{
  fftBase *f=((Pad *) pad)->fft;
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
  out[i++]=f->allRows();
  while(i < 32) out[i++]=0;
}

size_t fftwpp_pad_increment(void *pad, size_t r)
{
  return ((Pad *) pad)->fft->increment(r);
}
size_t fftwpp_pad_blocksize(void *pad, size_t r)
{
  return ((Pad *) pad)->fft->blocksize(r);
}
size_t fftwpp_pad_noutputs(void *pad, size_t r)
{
  return ((Pad *) pad)->fft->noutputs(r);
}
size_t fftwpp_pad_span(void *pad, size_t r)
{
  return ((Pad *) pad)->fft->span(r);
}
size_t fftwpp_pad_index(void *pad, size_t r, size_t i)
{
  return ((Pad *) pad)->fft->index(r,i);
}

void fftwpp_pad_forward(void *pad, const double *f, double *F, size_t r)
{
  ((Pad *) pad)->fft->forward((Complex *) f,(Complex *) F,r);
}

void fftwpp_pad_backward(void *pad, const double *F, double *f, size_t r)
{
  ((Pad *) pad)->fft->backward((Complex *) F,(Complex *) f,r);
}

// All-residues forward / backward pass with the OUTPUT rows split among
// `nsplit` owners of ceil-split row ranges (the slab decomposition's
// localdimension), every owner's range living in the same dense DEVICE buffer:
// the destination-set machinery of the fused exchange on one GPU.  The result
// must equal the plain all-residues pass.  Returns 0, or
// FFTWPP_GPU_EUNSUPPORTED when the plan has no TMA-staged kernel.
int fftwpp_pad_forward_split(void *pad, const void *f, void *F, size_t nsplit)
{
  fftBase *fft=((Pad *) pad)->fft;
  const std::vector<ResidueCall>& calls=fft->calls();
  size_t nsub=calls.back().sb0+calls.back().nsb;
  size_t rows=fft->allRows();
  std::vector<fftwpp_gpu_dest> d;
  for(size_t p=0; p < nsplit; ++p) {
    size_t r0;
    size_t n=utils::localdimension(rows,p,nsplit,&r0);
    fftwpp_gpu_dest t;
    t.base=(Complex *) F+r0*fft->S;
    t.row0=r0;
    t.rows=n;
    t.row_stride=fft->S;
    t.plane_stride=0;
    d.push_back(t);
  }
  return fftwpp_gpu_forward_dests(fft->plan(),0,nsub,f,d.data(),(int) d.size(),
                                  0,1,0,gpu::stream());
}

int fftwpp_pad_backward_split(void *pad, const void *F, void *f, size_t nsplit,
                              double scale)
{
  fftBase *fft=((Pad *) pad)->fft;
  const std::vector<ResidueCall>& calls=fft->calls();
  size_t nsub=calls.back().sb0+calls.back().nsb;
  size_t rows=fft->inputLength();
  std::vector<fftwpp_gpu_dest> d;
  for(size_t p=0; p < nsplit; ++p) {
    size_t r0;
    size_t n=utils::localdimension(rows,p,nsplit,&r0);
    fftwpp_gpu_dest t;
    t.base=(Complex *) f+r0*fft->S;
    t.row0=r0;
    t.rows=n;
    t.row_stride=fft->S;
    t.plane_stride=0;
    d.push_back(t);
  }
  return fftwpp_gpu_backward_dests(fft->plan(),0,nsub,F,d.data(),
                                   (int) d.size(),0,scale,1,0,gpu::stream());
}

void *fftwpp_conv_create(int dim, int family, const size_t *L, const size_t *M,
                         const size_t *m, const size_t *D, const long *I,
                         size_t Sx, size_t Sy, size_t A, size_t B, int mult)
{
  return makeConv(dim,family,L,M,m,D,I,Sx,Sy,A,B,pickMult(mult));
}

void *fftwpp_conv_create_custom(int dim, int family, const size_t *L,
                                const size_t *M, const size_t *m,
                                const size_t *D, const long *I, size_t Sx,
                                size_t Sy, size_t A, size_t B,
                                fftwpp_multiplier *host,
                                fftwpp_device_multiplier *device)
{
  if(!host) {
    std::cerr << "fftwpp_conv_create_custom: a host multiplier is required "
              << "(its address identifies the device one)" << std::endl;
    exit(-1);
  }
  registerDeviceMultiplier((multiplier *) host,(deviceMultiplier *) device);
  return makeConv(dim,family,L,M,m,D,I,Sx,Sy,A,B,(multiplier *) host);
}

void fftwpp_indices_get(void *indices, size_t *r, size_t *offset)
{
  Indices *i=(Indices *) indices;
  if(r) *r=i->r;
  if(offset) *offset=i->offset;
}

size_t fftwpp_indices_size(void *indices)
{
  return ((Indices *) indices)->size;
}

size_t fftwpp_indices_outer(void *indices, size_t d)
{
  Indices *i=(Indices *) indices;
  return d < i->size ? i->index[d] : 0;
}

size_t fftwpp_indices_index(void *indices, size_t j)
{
  Indices *i=(Indices *) indices;
  return i->fft->index(i->r,j+i->offset);
}

void fftwpp_conv_destroy(void *conv) {delete (Conv *) conv;}

void fftwpp_conv_params(void *conv, int d, size_t *out)
{
  fftBase *f=((Conv *) conv)->fft[d];
  out[0]=f->m; out[1]=f->p; out[2]=f->q; out[3]=f->n; out[4]=f->D;
  out[5]=f->inplace; out[6]=f->C; out[7]=f->S;
}

size_t fftwpp_conv_doubles(void *conv) {return ((Conv *) conv)->doubles;}

void fftwpp_conv_convolve(void *conv, double **f, int normalized)
{
  ((Conv *) conv)->convolve((Complex **) f,normalized != 0);
}

void fftwpp_conv_convolve_async(void *conv, double **f, int normalized,
                                int slot)
{
  Conv *c=(Conv *) conv;
  if(slot < 0 || slot > 1) {
    std::cerr << "fftwpp_conv_convolve_async: slot must be 0 or 1" << std::endl;
    exit(-1);
  }
  if(!c->h2d) {
    gpu::check(fftwpp_gpu_stream_create(&c->h2d),"stream creation");
    gpu::check(fftwpp_gpu_stream_create(&c->d2h),"stream creation");
    for(int s=0; s < 2; ++s) {
      gpu::check(fftwpp_gpu_event_create(&c->evIn[s]),"event creation");
      gpu::check(fftwpp_gpu_event_create(&c->evDone[s]),"event creation");
      gpu::check(fftwpp_gpu_event_create(&c->evOut[s]),"event creation");
    }
  }
  size_t N=std::max(c->A,c->B);
  size_t bytes=c->doubles*sizeof(double);
  c->slotBuf[slot].ensure(N,bytes);
  void *st=gpu::stream();
  // the slot's staging buffers are free once its previous outputs have left
  if(c->busy[slot])
    gpu::check(fftwpp_gpu_stream_wait_event(c->h2d,c->evOut[slot]),"wait");
  std::vector<Complex *> d(N);
  for(size_t a=0; a < N; ++a) d[a]=(Complex *) c->slotBuf[slot].ptr[a];
  for(size_t a=0; a < c->A; ++a)
    gpu::check(fftwpp_gpu_memcpy_h2d(d[a],f[a],bytes,c->h2d),"h2d");
  gpu::check(fftwpp_gpu_event_record(c->evIn[slot],c->h2d),"event");
  gpu::check(fftwpp_gpu_stream_wait_event(st,c->evIn[slot]),"wait");
  c->convolve(d.data(),normalized != 0);
  gpu::check(fftwpp_gpu_event_record(c->evDone[slot],st),"event");
  gpu::check(fftwpp_gpu_stream_wait_event(c->d2h,c->evDone[slot]),"wait");
  for(size_t b=0; b < c->B; ++b)
    gpu::check(fftwpp_gpu_memcpy_d2h(f[b],d[b],bytes,c->d2h),"d2h");
  gpu::check(fftwpp_gpu_event_record(c->evOut[slot],c->d2h),"event");
  c->busy[slot]=true;
}

void fftwpp_conv_wait(void *conv, int slot)
{
  Conv *c=(Conv *) conv;
  if(slot < 0 || slot > 1 || !c->busy[slot]) return;
  gpu::check(fftwpp_gpu_event_sync(c->evOut[slot]),"event sync");
  c->busy[slot]=false;
}

void fftwpp_conv_convolve_rows(void *conv, double **f, size_t nrows,
                               size_t rowstride, int normalized)
{
  Conv *c=(Conv *) conv;
  if(!c->c1) {
    std::cerr << "convolve_rows needs a 1-D convolution object" << std::endl;
    exit(-1);
  }
  c->c1->convolveRows((Complex **) f,0,nrows,rowstride,
                      normalized ? c->c1->scale : 1.0);
}

void fftwpp_conv_set_plane_chunk(void *conv, size_t chunk)
{
  Conv *c=(Conv *) conv;
  if(c->c3) c->c3->convolveyz[0]->planeChunk=chunk;
}

namespace {
struct MpiConv {
  int dim;
  Application *app[3];
  fftBase *fft[3];
  Convolution2MPI *conv2;
  Convolution3MPI *conv3;
  size_t A,B;
  size_t slabBytes; // bytes of one local input slab
  // pipelined host-buffer entry (fftwpp_mpiconv3_convolve_async)
  void *h2d,*d2h;
  void *evIn[2],*evDone[2],*evOut[2];
  bool busy[2];
  DeviceArrays slotBuf[2];
  MpiConv() : dim(0), conv2(NULL), conv3(NULL), A(0), B(0), slabBytes(0),
              h2d(NULL), d2h(NULL) {
    for(int d=0; d < 3; ++d) {app[d]=NULL; fft[d]=NULL;}
    busy[0]=busy[1]=false;
  }
  ~MpiConv() {
    if(h2d) {
      fftwpp_gpu_stream_destroy(h2d);
      fftwpp_gpu_stream_destroy(d2h);
      for(int s=0; s < 2; ++s) {
        fftwpp_gpu_event_destroy(evIn[s]);
        fftwpp_gpu_event_destroy(evDone[s]);
        fftwpp_gpu_event_destroy(evOut[s]);
      }
    }
    delete conv2;
    delete conv3;
    for(int d=2; d >= 0; --d) {delete fft[d]; delete app[d];}
  }
  void run(Complex **f, bool normalized) {
    if(conv3) {
      if(normalized) conv3->convolve(f);
      else conv3->convolveRaw(f);
    } else {
      if(normalized) conv2->convolve(f);
      else conv2->convolveRaw(f);
    }
  }
  SlabTranspose *slab() {
    return conv3 ? (SlabTranspose *) conv3 : (SlabTranspose *) conv2;
  }
};

// dim 2: arrays Lx x y (y split); dim 3: Lx x y x Lz.  family as in
// fftwpp_conv_create: 0 complex, 1 centred Hermitian (last dimension holds
// the ceil(L/2) non-negative modes), 2 real.
MpiConv *makeMpiConv(int dim, int family, const size_t *L, const size_t *M,
                     const size_t *m, const size_t *D, const long *I,
                     size_t A, size_t B, int mult, int rank, int size,
                     void *comm, int rankZ=0, int sizeZ=1, void *commZ=NULL)
{
  if(family < 0 || family > 2 || (dim == 2 && family == 2)) {
    std::cerr << "distributed convolutions: unsupported dim/family "
              << dim << "/" << family << std::endl;
    exit(-1);
  }
  size_t zero[3]={0,0,0};
  long minus[3]={-1,-1,-1};
  if(!m) m=zero;
  if(!D) D=zero;
  if(!I) I=minus;
  MpiConv *c=new MpiConv;
  c->dim=dim;
  utils::MPIgroup group(rank,size,comm);
  int kinds[3];
  size_t len[3];
  for(int d=0; d < dim; ++d) {
    if(family == 0) kinds[d]=0;
    else if(family == 1) kinds[d]=(d == dim-1) ? 2 : 1;
    else kinds[d]=(d == 0) ? 3 : 0;
    len[d]=(family == 1 && d == dim-1) ? ceilquotient(L[d],2) : L[d];
  }
  // the split dimension is y: in 2-D Hermitian runs that is the half-length
  // dimension, so the slices are taken over its stored modes
  size_t y0;
  size_t y=utils::localdimension(len[1],rank,size,&y0);
  // Ranks beyond the ceil-split own no y rows (reference localdimension,
  // mpi/mpitranspose.h:118-130, e.g. Ly=9 on 4 ranks: 3,3,3,0).  They still
  // own transformed x rows and take part in both exchanges; their local x
  // passes are skipped.  The plan of the (unused) x pass is built for one row.
  size_t yPlan=std::max<size_t>(y,1);
  // Every rank must run the x pass with the same (m,D,I): the reference picks
  // them on rank 0 and broadcasts (mpi/tests/hybridconvr3.cc:87-102); here the
  // deterministic chooser is evaluated for rank 0's slab width on every rank.
  size_t mx=m[0], Dx=D[0];
  long Ix=I[0];
  if(sizeZ > 1 && (dim != 3 || family == 1)) {
    std::cerr << "pencil decomposition: 3-D complex and real families only"
              << std::endl;
    exit(-1);
  }
  for(int d=0; d < dim; ++d) {
    multiplier *mu=(d == dim-1) ? pickMult(mult) : multNone;
    long Id=m[d] > 0 ? I[d] : -1;
    if(d == 0)
      c->app[d]=new Application(A,B,mu,fftw::maxthreads,false,m[d],D[d],Id);
    else
      c->app[d]=new Application(A,B,mu,*c->app[d-1],m[d],D[d],Id);
  }
  // pencil: only the z slice of the second group is local
  size_t rowWords=dim == 2 ? 1 :
    std::max<size_t>(utils::localdimension(len[2],rankZ,sizeZ,NULL),1);
  if(mx == 0) {
    size_t C0=utils::localdimension(len[1],0,size,NULL)*
      (dim == 2 ? 1 : utils::localdimension(len[2],0,sizeZ,NULL));
    fftBase *probe=makePad(kinds[0],L[0],M[0],*c->app[0],C0,C0,0,0,-1);
    mx=probe->m;
    Dx=probe->D;
    Ix=probe->inplace;
    delete probe;
  }
  if(dim == 2) {
    c->fft[0]=makePad(kinds[0],L[0],M[0],*c->app[0],yPlan,yPlan,mx,Dx,Ix);
    c->fft[1]=makePad(kinds[1],L[1],M[1],*c->app[1],1,0,m[1],D[1],I[1]);
    c->conv2=new Convolution2MPI(c->fft[0],c->fft[1],group);
  } else {
    size_t Cx=yPlan*rowWords;
    c->fft[0]=makePad(kinds[0],L[0],M[0],*c->app[0],Cx,Cx,mx,Dx,Ix);
    // every rank of the second group must run the y pass with the same
    // (m,D,I): chosen for the widest z slice (rank 0's)
    size_t my=m[1], Dy=D[1];
    long Iy=I[1];
    if(sizeZ > 1 && my == 0) {
      size_t z0=utils::localdimension(len[2],0,sizeZ,NULL);
      fftBase *probe=makePad(kinds[1],L[1],M[1],*c->app[1],z0,z0,0,0,-1);
      my=probe->m;
      Dy=probe->D;
      Iy=probe->inplace;
      delete probe;
    }
    c->fft[1]=makePad(kinds[1],L[1],M[1],*c->app[1],rowWords,rowWords,my,Dy,
                      Iy);
    c->fft[2]=makePad(kinds[2],L[2],M[2],*c->app[2],1,0,m[2],D[2],I[2]);
    if(sizeZ > 1) {
      utils::MPIgroup groupZ(rankZ,sizeZ,commZ);
      c->conv3=new Convolution3MPI(c->fft[0],c->fft[1],c->fft[2],group,groupZ);
    } else
      c->conv3=new Convolution3MPI(c->fft[0],c->fft[1],c->fft[2],group);
  }
  c->A=A;
  c->B=B;
  c->slabBytes=L[0]*y*rowWords*(family == 2 ? sizeof(double) : sizeof(Complex));
  return c;
}
}

void *fftwpp_mpiconv3_create(int family, const size_t *L, const size_t *M,
                             const size_t *m, const size_t *D, const long *I,
                             size_t A, size_t B, int mult, int rank, int size,
                             void *comm)
{
  return makeMpiConv(3,family,L,M,m,D,I,A,B,mult,rank,size,comm);
}

// Pencil decomposition (forced test / bench mode on one box): y split over
// the first group, z over the second; arrays are the local pencils
// Lx x y x z.  out of fftwpp_mpiconv3_split then describes the first group's
// exchange with Z = the local z extent.
void *fftwpp_mpiconv3_create_pencil(int family, const size_t *L,
                                    const size_t *M, const size_t *m,
                                    const size_t *D, const long *I, size_t A,
                                    size_t B, int mult, int rankY, int sizeY,
                                    void *commY, int rankZ, int sizeZ,
                                    void *commZ)
{
  return makeMpiConv(3,family,L,M,m,D,I,A,B,mult,rankY,sizeY,commY,rankZ,sizeZ,
                     commZ);
}

void *fftwpp_mpiconv2_create(int family, const size_t *L, const size_t *M,
                             const size_t *m, const size_t *D, const long *I,
                             size_t A, size_t B, int mult, int rank, int size,
                             void *comm)
{
  return makeMpiConv(2,family,L,M,m,D,I,A,B,mult,rank,size,comm);
}

void fftwpp_mpiconv3_destroy(void *conv) {delete (MpiConv *) conv;}
void fftwpp_mpiconv2_destroy(void *conv) {delete (MpiConv *) conv;}

void fftwpp_mpiconv3_split(void *conv, size_t *out)
{
  utils::split3& d=((MpiConv *) conv)->slab()->d;
  out[0]=d.X; out[1]=d.Y; out[2]=d.Z; out[3]=d.x; out[4]=d.y; out[5]=d.z;
  out[6]=d.x0; out[7]=d.y0; out[8]=d.z0;
}
void fftwpp_mpiconv2_split(void *conv, size_t *out)
{
  fftwpp_mpiconv3_split(conv,out);
}

void fftwpp_mpiconv3_params(void *conv, int d, size_t *out)
{
  fftBase *f=((MpiConv *) conv)->fft[d];
  out[0]=f->m; out[1]=f->p; out[2]=f->q; out[3]=f->n; out[4]=f->D;
  out[5]=f->inplace; out[6]=f->C; out[7]=f->S;
}
void fftwpp_mpiconv2_params(void *conv, int d, size_t *out)
{
  fftwpp_mpiconv3_params(conv,d,out);
}

void fftwpp_mpiconv3_convolve(void *conv, double **f, int normalized)
{
  ((MpiConv *) conv)->run((Complex **) f,normalized != 0);
}

// Pipelined form for PINNED HOST slabs: every rank copies its own slabs over
// its own PCIe link on a side stream, convolves on the compute stream (the
// exchanges are stream-ordered collectives there) and copies the result back
// on a third stream.  Two slots: alternate them so that one convolution's
// transfers overlap the other's compute.  COLLECTIVE: all ranks must call it
// in the same order.
void fftwpp_mpiconv3_convolve_async(void *conv, double **f, int normalized,
                                    int slot)
{
  MpiConv *c=(MpiConv *) conv;
  if(slot < 0 || slot > 1) {
    std::cerr << "fftwpp_mpiconv3_convolve_async: slot must be 0 or 1"
              << std::endl;
    exit(-1);
  }
  if(!c->h2d) {
    gpu::check(fftwpp_gpu_stream_create(&c->h2d),"stream creation");
    gpu::check(fftwpp_gpu_stream_create(&c->d2h),"stream creation");
    for(int s=0; s < 2; ++s) {
      gpu::check(fftwpp_gpu_event_create(&c->evIn[s]),"event creation");
      gpu::check(fftwpp_gpu_event_create(&c->evDone[s]),"event creation");
      gpu::check(fftwpp_gpu_event_create(&c->evOut[s]),"event creation");
    }
  }
  size_t N=std::max(c->A,c->B);
  size_t bytes=c->slabBytes;
  c->slotBuf[slot].ensure(N,std::max<size_t>(bytes,16));
  void *st=gpu::stream();
  if(c->busy[slot])
    gpu::check(fftwpp_gpu_stream_wait_event(c->h2d,c->evOut[slot]),"wait");
  std::vector<Complex *> d(N);
  for(size_t a=0; a < N; ++a) d[a]=(Complex *) c->slotBuf[slot].ptr[a];
  for(size_t a=0; a < c->A && bytes; ++a)
    gpu::check(fftwpp_gpu_memcpy_h2d(d[a],f[a],bytes,c->h2d),"h2d");
  gpu::check(fftwpp_gpu_event_record(c->evIn[slot],c->h2d),"event");
  gpu::check(fftwpp_gpu_stream_wait_event(st,c->evIn[slot]),"wait");
  c->run(d.data(),normalized != 0);
  gpu::check(fftwpp_gpu_event_record(c->evDone[slot],st),"event");
  gpu::check(fftwpp_gpu_stream_wait_event(c->d2h,c->evDone[slot]),"wait");
  for(size_t b=0; b < c->B && bytes; ++b)
    gpu::check(fftwpp_gpu_memcpy_d2h(f[b],d[b],bytes,c->d2h),"d2h");
  gpu::check(fftwpp_gpu_event_record(c->evOut[slot],c->d2h),"event");
  c->busy[slot]=true;
}

void fftwpp_mpiconv3_wait(void *conv, int slot)
{
  MpiConv *c=(MpiConv *) conv;
  if(slot < 0 || slot > 1 || !c->busy[slot]) return;
  gpu::check(fftwpp_gpu_event_sync(c->evOut[slot]),"event sync");
  c->busy[slot]=false;
}
void fftwpp_mpiconv2_convolve(void *conv, double **f, int normalized)
{
  fftwpp_mpiconv3_convolve(conv,f,normalized);
}

void fftwpp_mpiconv3_exchange_table(void *conv, int direction,
                                    unsigned long long *scount,
                                    unsigned long long *sdispl,
                                    unsigned long long *rcount,
                                    unsigned long long *rdispl)
{
  ((MpiConv *) conv)->slab()->exchangeTable(direction,(uint64_t *) scount,
                                            (uint64_t *) sdispl,
                                            (uint64_t *) rcount,
                                            (uint64_t *) rdispl);
}
void fftwpp_mpiconv2_exchange_table(void *conv, int direction,
                                    unsigned long long *scount,
                                    unsigned long long *sdispl,
                                    unsigned long long *rcount,
                                    unsigned long long *rdispl)
{
  fftwpp_mpiconv3_exchange_table(conv,direction,scount,sdispl,rcount,rdispl);
}

void fftwpp_mpiconv3_symmetrize(void *conv, double *f)
{
  MpiConv *c=(MpiConv *) conv;
  if(!c->conv3) {
    std::cerr << "fftwpp_mpiconv3_symmetrize needs a 3-D handle" << std::endl;
    exit(-1);
  }
  c->conv3->HermitianSymmetrizeXY((Complex *) f);
}

void fftwpp_mpiconv3_set_plane_chunk(void *conv, size_t chunk)
{
  MpiConv *c=(MpiConv *) conv;
  if(c->conv3) c->conv3->convolveyz[0]->planeChunk=chunk;
}

// ---- distributed FFTs (cpp/mpifftw++.h) ----
namespace {
struct MpiFft {
  int kind,dims;
  utils::MPIgroup group;
  fft2dMPI *c;
  rcfft2dMPI *r;
  MpiFft(int rank, int size, void *comm) : group(rank,size,comm), c(NULL),
                                           r(NULL) {}
  ~MpiFft() {delete c; delete r;}
  fftMPIBase *base() {return c ? (fftMPIBase *) c : (fftMPIBase *) r;}
};
}

void *fftwpp_mpifft_create(int kind, int dims, const size_t *N, int sign,
                           int rank, int size, void *comm)
{
  if((kind != 0 && kind != 1) || (dims != 2 && dims != 3)) {
    std::cerr << "fftwpp_mpifft_create: kind must be 0 or 1, dims 2 or 3"
              << std::endl;
    exit(-1);
  }
  MpiFft *h=new MpiFft(rank,size,comm);
  h->kind=kind;
  h->dims=dims;
  if(kind == 0) {
    if(dims == 2)
      h->c=new fft2dMPI(utils::split(N[0],N[1],h->group),h->group,sign);
    else
      h->c=new fft3dMPI(utils::split3(N[0],N[1],N[2],h->group),h->group,sign);
  } else {
    if(dims == 2)
      h->r=new rcfft2dMPI(utils::split(N[0],N[1],h->group),
                          utils::split(N[0],N[1]/2+1,h->group),h->group);
    else
      h->r=new rcfft3dMPI(utils::split3(N[0],N[1],N[2],h->group),
                          utils::split3(N[0],N[1],N[2]/2+1,h->group),
                          h->group);
  }
  return h;
}

void fftwpp_mpifft_destroy(void *fft) {delete (MpiFft *) fft;}

void fftwpp_mpifft_split(void *fft, size_t *out)
{
  utils::split3& d=((MpiFft *) fft)->base()->d;
  out[0]=d.X; out[1]=d.Y; out[2]=d.Z; out[3]=d.x; out[4]=d.y; out[5]=d.z;
  out[6]=d.x0; out[7]=d.y0; out[8]=d.z0;
}

size_t fftwpp_mpifft_words(void *fft) {return ((MpiFft *) fft)->base()->n();}

void fftwpp_mpifft_exchange_table(void *fft, int direction,
                                  unsigned long long *scount,
                                  unsigned long long *sdispl,
                                  unsigned long long *rcount,
                                  unsigned long long *rdispl)
{
  ((MpiFft *) fft)->base()->exchangeTable(direction,(uint64_t *) scount,
                                          (uint64_t *) sdispl,
                                          (uint64_t *) rcount,
                                          (uint64_t *) rdispl);
}

void fftwpp_mpifft_forward(void *fft, void *in, void *out)
{
  MpiFft *h=(MpiFft *) fft;
  if(h->c) h->c->Forward((Complex *) in,(Complex *) out);
  else h->r->Forward((double *) in,(Complex *) out);
}

void fftwpp_mpifft_backward(void *fft, void *in, void *out)
{
  MpiFft *h=(MpiFft *) fft;
  if(h->c) h->c->Backward((Complex *) in,(Complex *) out);
  else h->r->Backward((Complex *) in,(double *) out);
}

void fftwpp_mpifft_shift(void *fft, double *f)
{
  MpiFft *h=(MpiFft *) fft;
  if(!h->r) {
    std::cerr << "fftwpp_mpifft_shift needs a real-to-complex handle"
              << std::endl;
    exit(-1);
  }
  h->r->Shift(f);
}

void fftwpp_mpifft_denyquist(void *fft, void *f)
{
  MpiFft *h=(MpiFft *) fft;
  if(!h->r) {
    std::cerr << "fftwpp_mpifft_denyquist needs a real-to-complex handle"
              << std::endl;
    exit(-1);
  }
  h->r->deNyquist((Complex *) f);
}

void fftwpp_mpifft_normalize(void *fft, void *f)
{
  MpiFft *h=(MpiFft *) fft;
  if(h->c) h->c->Normalize((Complex *) f);
  else h->r->Normalize((double *) f);
}

void fftwpp_set_stream(void *stream) {gpu::setStream(stream);}

}
