// mpiconvolve.cc -- see mpiconvolve.h.
#include "mpiconvolve.h"
#include "../../include/fftwpp_gpu.h"

#include <algorithm>
#include <vector>

using namespace utils;

namespace fftwpp {

Convolution3MPI::Convolution3MPI(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                                 const MPIgroup& group) :
  Convolution3(fftx,ffty,fftz,NULL,NULL,NULL,true), group(group)
{
  if(fftx->S != fftx->C) {
    std::cerr << "Convolution3MPI: the local x pass must be contiguous (S == C)"
              << std::endl;
    exit(-1);
  }
  size_t rowWords=ffty->S;          // words per y row (z extent incl. stride)
  d=split3(fftx->allRows(),ffty->L,rowWords,group);
  if(fftx->C != d.y*rowWords) {
    std::cerr << "Convolution3MPI: fftx->C=" << fftx->C
              << " does not match the local slab " << d.y << "x" << rowWords
              << std::endl;
    exit(-1);
  }
  delete convolveyz[0];
  convolveyz[0]=new Convolution2(ffty,fftz);
  fftx->setTag(1);
  ffty->setTag(2);
  fftz->setTag(3);
  scale=1.0/normalization();
}

Convolution3MPI::~Convolution3MPI() {}

// direction 0: x-transformed slab (X x y x Z) -> (x x Y x Z)   [localize1]
// direction 1: the inverse                                     [localize0]
void Convolution3MPI::exchangeTable(int direction, uint64_t *scount,
                                    uint64_t *sdispl, uint64_t *rcount,
                                    uint64_t *rdispl)
{
  const uint64_t w=sizeof(Complex);
  uint64_t soff=0, roff=0;
  for(int p=0; p < group.size; ++p) {
    size_t px0,py0;
    size_t px=localdimension(d.X,p,group.size,&px0);
    size_t py=localdimension(d.Y,p,group.size,&py0);
    // forward: to p go its x rows of my y slice (contiguous rows of Fx);
    // from p come my x rows of p's y slice (packed per source)
    uint64_t toP=(uint64_t) px*d.y*d.Z*w;
    uint64_t fromP=(uint64_t) d.x*py*d.Z*w;
    if(direction == 0) {
      scount[p]=toP;
      sdispl[p]=(uint64_t) px0*d.y*d.Z*w;
      rcount[p]=fromP;
      rdispl[p]=roff;
      roff += fromP;
    } else {
      scount[p]=fromP;
      sdispl[p]=soff;
      soff += fromP;
      rcount[p]=toP;
      rdispl[p]=(uint64_t) px0*d.y*d.Z*w;
    }
  }
}

void Convolution3MPI::transposeForward(void *Fx, void *T)
{
  std::vector<uint64_t> sc(group.size),sd(group.size),rc(group.size),
    rd(group.size);
  exchangeTable(0,sc.data(),sd.data(),rc.data(),rd.data());
  void *st=gpu::stream();
  gpu::check(fftwpp_gpu_comm_alltoallv(group.comm,Fx,sc.data(),sd.data(),
                                       devP.ptr[0],rc.data(),rd.data(),st),
             "all-to-all (localize1)");
  // unpack: block from p holds [x][py][Z] -> T[x][py0+..][Z]
  for(int p=0; p < group.size; ++p) {
    size_t py0;
    size_t py=localdimension(d.Y,p,group.size,&py0);
    if(py == 0 || d.x == 0) continue;
    gpu::check(fftwpp_gpu_copy3((Complex *) T+py0*d.Z,
                                (const char *) devP.ptr[0]+rd[p],
                                d.x,py,d.Z,d.Y*d.Z,d.Z,py*d.Z,d.Z,st),
               "unpack");
  }
}

void Convolution3MPI::transposeBackward(void *T, void *Fx)
{
  std::vector<uint64_t> sc(group.size),sd(group.size),rc(group.size),
    rd(group.size);
  exchangeTable(1,sc.data(),sd.data(),rc.data(),rd.data());
  void *st=gpu::stream();
  for(int p=0; p < group.size; ++p) {
    size_t py0;
    size_t py=localdimension(d.Y,p,group.size,&py0);
    if(py == 0 || d.x == 0) continue;
    gpu::check(fftwpp_gpu_copy3((char *) devP.ptr[0]+sd[p],
                                (const Complex *) T+py0*d.Z,
                                d.x,py,d.Z,py*d.Z,d.Z,d.Y*d.Z,d.Z,st),
               "pack");
  }
  gpu::check(fftwpp_gpu_comm_alltoallv(group.comm,devP.ptr[0],sc.data(),
                                       sd.data(),Fx,rc.data(),rd.data(),st),
             "all-to-all (localize0)");
}

void Convolution3MPI::runMPI(Complex **f, size_t offset, double sc)
{
  size_t N=std::max(A,B);
  void *st=gpu::stream();
  if(!gpu::isDevice(f[0])) {
    std::cerr << "Convolution3MPI needs device pointers" << std::endl;
    exit(-1);
  }
  size_t rows=fftx->allRows();
  size_t slabWords=rows*fftx->S;               // X x y x Z
  size_t tWords=std::max<size_t>(d.x,1)*d.Y*d.Z;
  devF.ensure(N,slabWords*sizeof(Complex));
  devT.ensure(N,tWords*sizeof(Complex));
  devP.ensure(1,std::max(slabWords,tWords)*sizeof(Complex));
  const std::vector<ResidueCall>& calls=fftx->calls();
  size_t nsub=calls.back().sb0+calls.back().nsb;

  std::vector<Complex *> T(N);
  for(size_t a=0; a < N; ++a) T[a]=(Complex *) devT.ptr[a];

  for(size_t a=0; a < A; ++a) {
    gpu::check(fftwpp_gpu_forward(fftx->plan(),0,nsub,1,f[a]+offset,
                                  devF.ptr[a],1,0,0,st),"forward");
    transposeForward(devF.ptr[a],devT.ptr[a]);
  }
  if(d.x > 0)
    convolveyz[0]->convolvePlanes(T.data(),0,d.x,d.Y*d.Z,1.0);
  for(size_t b=0; b < B; ++b) {
    transposeBackward(devT.ptr[b],devF.ptr[b]);
    gpu::check(fftwpp_gpu_backward(fftx->plan(),0,nsub,1,devF.ptr[b],
                                   f[b]+offset,0,sc,1,0,0,st),"backward");
  }
}

void Convolution3MPI::convolveRaw(Complex **f, size_t offset, Indices *)
{
  runMPI(f,offset,1.0);
}

void Convolution3MPI::convolve(Complex **f, size_t offset)
{
  runMPI(f,offset,scale);
}

} // namespace fftwpp
