// mpiconvolve.cc -- see mpiconvolve.h.
#include "mpiconvolve.h"
#include "../../include/fftwpp_gpu.h"

#include <algorithm>
#include <vector>

using namespace utils;

namespace fftwpp {

SlabTranspose::SlabTranspose(const MPIgroup& group) : group(group)
{
  emptyZ=false;
  commStream=NULL;
  nchunks=1;
  const char *e=getenv("FFTWPP_MPI_CHUNKS");
  if(e && *e) nchunks=std::max<size_t>(1,strtoull(e,NULL,10));
}

SlabTranspose::~SlabTranspose()
{
  for(size_t i=0; i < events.size(); ++i) fftwpp_gpu_event_destroy(events[i]);
  if(commStream) fftwpp_gpu_stream_destroy(commStream);
}

Convolution2MPI::Convolution2MPI(fftBase *fftx, fftBase *ffty,
                                 const MPIgroup& group) :
  Convolution2(fftx,ffty), SlabTranspose(group)
{
  if(fftx->S != fftx->C) {
    std::cerr << "Convolution2MPI: the local x pass must be contiguous (S == C)"
              << std::endl;
    exit(-1);
  }
  d=split3(fftx->allRows(),ffty->inputLength(),1,group);
  if(fftx->C != std::max<size_t>(d.y,1)) {
    std::cerr << "Convolution2MPI: fftx->C=" << fftx->C
              << " does not match the local slab width " << d.y << std::endl;
    exit(-1);
  }
  fftx->setTag(1);
  ffty->setTag(2);
  scale=1.0/normalization();
}

void Convolution2MPI::runMPI(Complex **f, size_t offset, double sc)
{
  runSlab(fftx,A,B,devF,f,offset,sc,[this](Complex **T, size_t lo, size_t hi) {
    convolvey[0]->convolveRows(T,lo*d.Y,hi-lo,d.Y,1.0);
  });
}

void Convolution2MPI::convolveRaw(Complex **f, size_t offset, Indices *)
{
  runMPI(f,offset,1.0);
}

void Convolution2MPI::convolve(Complex **f, size_t offset)
{
  runMPI(f,offset,scale);
}

Convolution3MPI::Convolution3MPI(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                                 const MPIgroup& group) :
  Convolution3(fftx,ffty,fftz,NULL,NULL,NULL,true), SlabTranspose(group),
  inner(NULL)
{
  init(group,ffty->S);
}

Convolution3MPI::Convolution3MPI(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                                 const MPIgroup& group,
                                 const MPIgroup& groupYZ) :
  Convolution3(fftx,ffty,fftz,NULL,NULL,NULL,true), SlabTranspose(group),
  inner(NULL)
{
  if(groupYZ.size > 1) {
    inner=new BatchedTranspose(groupYZ,ffty->allRows(),fftz->inputLength());
    emptyZ=inner->z == 0; // more ranks than z words: this pencil is empty
    if(ffty->C != std::max<size_t>(inner->z,1) || ffty->S != ffty->C) {
      std::cerr << "Convolution3MPI (pencil): ffty must be built for the "
                << "local z slice (C = S = " << inner->z << ")" << std::endl;
      exit(-1);
    }
  }
  init(group,ffty->S);
}

// zlocal: words per y row of the local data (all of z for slabs, the z slice
// of a pencil)
void Convolution3MPI::init(const MPIgroup& group, size_t zlocal)
{
  if(fftx->S != fftx->C) {
    std::cerr << "Convolution3MPI: the local x pass must be contiguous (S == C)"
              << std::endl;
    exit(-1);
  }
  size_t rowWords=zlocal;          // words per y row (z extent incl. stride)
  d=split3(fftx->allRows(),ffty->L,rowWords,group);
  if(fftx->C != std::max<size_t>(d.y,1)*rowWords) {
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
  // the fused exchange needs both strided passes on the power-of-two
  // register kernels (fast_kernels.cu) and the y pass in its direct variant
  auto pow2ok=[](size_t m) {return m >= 64 && m <= 4096 && (m & (m-1)) == 0;};
  fused=!inner && pow2ok(fftx->m) && pow2ok(ffty->m) && ffty->p == 1 &&
    ffty->kind() == fftBase::COMPLEX && ffty->C >= 4 && fftx->C >= 4 &&
    fftx->q > 1 && ffty->q > 1 &&
    (fftx->kind() == fftBase::COMPLEX || fftx->kind() == fftBase::REAL) &&
    fftx->p <= 2;
  const char *e=getenv("FFTWPP_MPI_FUSED");
  if(e && *e == '0') fused=false;
  // kernel eligibility (fast plans present, register variant of the y pass)
  // -- probed only where a communicator (hence a device) exists
  if(fused && group.comm && d.y > 0)
    fused=fftwpp_gpu_mapped_supported(fftx->plan(),0) == 1 &&
      fftwpp_gpu_mapped_supported(ffty->plan(),1) == 1;
  // The predicate depends on rank-local values (slab width, environment):
  // all ranks must take the same path or they would wait in different
  // collectives.  Reduce it (logical AND) over the group.
  fused=agree(fused);
  fusedReady=false;
}

// Logical AND of a rank-local predicate over all ranks (collective).
bool SlabTranspose::agree(bool mine)
{
  if(!group.comm || group.size <= 1) return mine;
  void *st=gpu::stream();
  DeviceArrays tmp;
  tmp.ensure(2,sizeof(uint64_t)*group.size);
  uint64_t v=mine ? 1 : 0;
  std::vector<uint64_t> all(group.size);
  gpu::check(fftwpp_gpu_memcpy_h2d(tmp.ptr[0],&v,sizeof(v),st),"h2d");
  gpu::check(fftwpp_gpu_comm_allgather(group.comm,tmp.ptr[0],tmp.ptr[1],
                                       sizeof(v),st),"all-gather (agree)");
  gpu::check(fftwpp_gpu_memcpy_d2h(all.data(),tmp.ptr[1],
                                   sizeof(v)*group.size,st),"d2h");
  gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  bool ok=true;
  for(int p=0; p < group.size; ++p) ok=ok && all[p] != 0;
  return ok;
}

Convolution3MPI::~Convolution3MPI()
{
  for(size_t i=0; i < opened.size(); ++i) fftwpp_gpu_ipc_close(opened[i]);
  delete inner;
}

// ---------------------------------------------------------------------------
// pencil decomposition
// ---------------------------------------------------------------------------

BatchedTranspose::BatchedTranspose(const MPIgroup& group, size_t R, size_t Z) :
  group(group), R(R), Z(Z)
{
  r=localdimension(R,group.rank,group.size,&r0);
  z=localdimension(Z,group.rank,group.size,&z0);
}

void BatchedTranspose::forward(const void *U, void *V, size_t planes,
                               DeviceArrays& scratch, void *st)
{
  const int P=group.size;
  const uint64_t w=sizeof(Complex);
  std::vector<uint64_t> sc(P),sd(P),rc(P),rd(P);
  uint64_t so=0,ro=0;
  for(int p=0; p < P; ++p) {
    size_t pr=localdimension(R,p,P,NULL);
    size_t pz=localdimension(Z,p,P,NULL);
    sc[p]=(uint64_t) planes*pr*z*w;   // to p: its rows of my z slice
    sd[p]=so;
    so += sc[p];
    rc[p]=(uint64_t) planes*r*pz*w;   // from p: my rows of its z slice
    rd[p]=ro;
    ro += rc[p];
  }
  scratch.ensure(2,std::max<uint64_t>(std::max(so,ro),16));
  char *sendb=(char *) scratch.ptr[0];
  char *recvb=(char *) scratch.ptr[1];
  for(int p=0; p < P; ++p) {
    size_t pr0;
    size_t pr=localdimension(R,p,P,&pr0);
    if(pr == 0 || z == 0 || planes == 0) continue;
    // [plane][row in p's range][k] <- U[plane][pr0+row][k]
    gpu::check(fftwpp_gpu_copy3(sendb+sd[p],(const Complex *) U+pr0*z,planes,
                                pr,z,pr*z,z,R*z,z,st),"pack (pencil)");
  }
  gpu::check(fftwpp_gpu_comm_alltoallv(group.comm,sendb,sc.data(),sd.data(),
                                       recvb,rc.data(),rd.data(),st),
             "all-to-all (pencil, forward)");
  for(int p=0; p < P; ++p) {
    size_t pz0;
    size_t pz=localdimension(Z,p,P,&pz0);
    if(pz == 0 || r == 0 || planes == 0) continue;
    gpu::check(fftwpp_gpu_copy3((Complex *) V+pz0,recvb+rd[p],planes,r,pz,r*Z,
                                Z,r*pz,pz,st),"unpack (pencil)");
  }
}

void BatchedTranspose::backward(const void *V, void *U, size_t planes,
                                DeviceArrays& scratch, void *st)
{
  const int P=group.size;
  const uint64_t w=sizeof(Complex);
  std::vector<uint64_t> sc(P),sd(P),rc(P),rd(P);
  uint64_t so=0,ro=0;
  for(int p=0; p < P; ++p) {
    size_t pr=localdimension(R,p,P,NULL);
    size_t pz=localdimension(Z,p,P,NULL);
    sc[p]=(uint64_t) planes*r*pz*w;   // to p: my rows of its z slice
    sd[p]=so;
    so += sc[p];
    rc[p]=(uint64_t) planes*pr*z*w;   // from p: its rows of my z slice
    rd[p]=ro;
    ro += rc[p];
  }
  scratch.ensure(2,std::max<uint64_t>(std::max(so,ro),16));
  char *sendb=(char *) scratch.ptr[0];
  char *recvb=(char *) scratch.ptr[1];
  for(int p=0; p < P; ++p) {
    size_t pz0;
    size_t pz=localdimension(Z,p,P,&pz0);
    if(pz == 0 || r == 0 || planes == 0) continue;
    gpu::check(fftwpp_gpu_copy3(sendb+sd[p],(const Complex *) V+pz0,planes,r,
                                pz,r*pz,pz,r*Z,Z,st),"pack (pencil)");
  }
  gpu::check(fftwpp_gpu_comm_alltoallv(group.comm,sendb,sc.data(),sd.data(),
                                       recvb,rc.data(),rd.data(),st),
             "all-to-all (pencil, backward)");
  for(int p=0; p < P; ++p) {
    size_t pr0;
    size_t pr=localdimension(R,p,P,&pr0);
    if(pr == 0 || z == 0 || planes == 0) continue;
    gpu::check(fftwpp_gpu_copy3((Complex *) U+pr0*z,recvb+rd[p],planes,pr,z,
                                R*z,z,pr*z,z,st),"unpack (pencil)");
  }
}

// x pass local -> xy exchange over `group` (blocks of the local z slice) ->
// for the local transformed x rows: y pass local, yz exchange over the second
// group, z convolutions on rows with all of z, inverse yz exchange, y backward
// -> inverse xy exchange -> x backward  (reference mpi/mpiconvolve.h:208-305)
void Convolution3MPI::runPencil(Complex **f, size_t offset, double sc)
{
  runSlab(fftx,A,B,devF,f,offset,sc,[this](Complex **T, size_t lo, size_t hi) {
    size_t N=std::max(A,B);
    size_t planes=hi-lo;
    void *st=gpu::stream();
    size_t zl=inner->z;                 // local z words
    size_t R=inner->R;                  // transformed y rows
    size_t planeIn=d.Y*zl;              // words per plane of T
    const std::vector<ResidueCall>& calls=ffty->calls();
    size_t nsub=calls.back().sb0+calls.back().nsb;
    devU.ensure(N,std::max<size_t>(planes*R*zl,1)*sizeof(Complex));
    devV.ensure(N,std::max<size_t>(planes*inner->r*inner->Z,1)*sizeof(Complex));
    std::vector<Complex *> V(N);
    for(size_t a=0; a < N; ++a) V[a]=(Complex *) devV.ptr[a];
    for(size_t a=0; a < A; ++a) {
      if(zl > 0)
        gpu::check(fftwpp_gpu_forward(ffty->plan(),0,nsub,1,
                                      T[a]+lo*planeIn,devU.ptr[a],planes,
                                      planeIn,R*zl,st),"forward (pencil y)");
      inner->forward(devU.ptr[a],devV.ptr[a],planes,devS,st);
    }
    if(inner->r > 0)
      convolveyz[0]->convolvey[0]->convolveRows(V.data(),0,planes*inner->r,inner->Z,1.0);
    for(size_t b=0; b < B; ++b) {
      inner->backward(devV.ptr[b],devU.ptr[b],planes,devS,st);
      if(zl > 0)
        gpu::check(fftwpp_gpu_backward(ffty->plan(),0,nsub,1,devU.ptr[b],
                                       T[b]+lo*planeIn,0,1.0,planes,R*zl,
                                       planeIn,st),"backward (pencil y)");
    }
  });
}

// Sub-range [lo,hi) (relative to the rank's first transformed x row) of
// chunk c of nc for the given rank.
void SlabTranspose::chunkRange(int rank, size_t c, size_t nc, size_t *lo,
                                 size_t *hi)
{
  size_t x=localdimension(d.X,rank,group.size,NULL);
  *lo=c*x/nc;
  *hi=(c+1)*x/nc;
}

// direction 0: x-transformed slab (X x y x Z) -> (x x Y x Z)   [localize1]
// direction 1: the inverse                                     [localize0]
void SlabTranspose::exchangeTable(int direction, uint64_t *scount,
                                    uint64_t *sdispl, uint64_t *rcount,
                                    uint64_t *rdispl, size_t c, size_t nc)
{
  const uint64_t w=sizeof(Complex);
  uint64_t off=0;
  size_t mylo,myhi;
  chunkRange(group.rank,c,nc,&mylo,&myhi);
  for(int p=0; p < group.size; ++p) {
    size_t px0,py0,plo,phi;
    localdimension(d.X,p,group.size,&px0);
    size_t py=localdimension(d.Y,p,group.size,&py0);
    chunkRange(p,c,nc,&plo,&phi);
    // forward: to p go its chunk rows of my y slice (contiguous rows of Fx);
    // from p come my chunk rows of p's y slice (packed per source)
    uint64_t toP=(uint64_t) (phi-plo)*d.y*d.Z*w;
    uint64_t fromP=(uint64_t) (myhi-mylo)*py*d.Z*w;
    uint64_t slabOff=(uint64_t) (px0+plo)*d.y*d.Z*w;
    if(direction == 0) {
      scount[p]=toP;
      sdispl[p]=slabOff;
      rcount[p]=fromP;
      rdispl[p]=off;
    } else {
      scount[p]=fromP;
      sdispl[p]=off;
      rcount[p]=toP;
      rdispl[p]=slabOff;
    }
    off += fromP;
  }
}

void SlabTranspose::transposeForward(void *Fx, void *T, size_t c, size_t nc,
                                       void *st)
{
  std::vector<uint64_t> sc(group.size),sd(group.size),rc(group.size),
    rd(group.size);
  exchangeTable(0,sc.data(),sd.data(),rc.data(),rd.data(),c,nc);
  size_t lo,hi;
  chunkRange(group.rank,c,nc,&lo,&hi);
  gpu::check(fftwpp_gpu_comm_alltoallv(group.comm,Fx,sc.data(),sd.data(),
                                       devP.ptr[0],rc.data(),rd.data(),st),
             "all-to-all (localize1)");
  // unpack: block from p holds [rows][py][Z] -> T[lo+..][py0+..][Z]
  for(int p=0; p < group.size; ++p) {
    size_t py0;
    size_t py=localdimension(d.Y,p,group.size,&py0);
    if(py == 0 || hi == lo) continue;
    gpu::check(fftwpp_gpu_copy3((Complex *) T+lo*d.Y*d.Z+py0*d.Z,
                                (const char *) devP.ptr[0]+rd[p],
                                hi-lo,py,d.Z,d.Y*d.Z,d.Z,py*d.Z,d.Z,st),
               "unpack");
  }
}

void SlabTranspose::transposeBackward(void *T, void *Fx, size_t c, size_t nc,
                                        void *st)
{
  std::vector<uint64_t> sc(group.size),sd(group.size),rc(group.size),
    rd(group.size);
  exchangeTable(1,sc.data(),sd.data(),rc.data(),rd.data(),c,nc);
  size_t lo,hi;
  chunkRange(group.rank,c,nc,&lo,&hi);
  for(int p=0; p < group.size; ++p) {
    size_t py0;
    size_t py=localdimension(d.Y,p,group.size,&py0);
    if(py == 0 || hi == lo) continue;
    gpu::check(fftwpp_gpu_copy3((char *) devP.ptr[0]+sd[p],
                                (const Complex *) T+lo*d.Y*d.Z+py0*d.Z,
                                hi-lo,py,d.Z,py*d.Z,d.Z,d.Y*d.Z,d.Z,st),
               "pack");
  }
  gpu::check(fftwpp_gpu_comm_alltoallv(group.comm,devP.ptr[0],sc.data(),
                                       sd.data(),Fx,rc.data(),rd.data(),st),
             "all-to-all (localize0)");
}

// Pipeline (compute stream C = gpu::stream(), exchange stream X):
//   C: xfwd(0) xfwd(1) .. | yz(chunk 0) | yz(chunk 1) | ...        | xbwd
//   X:        F(0,c0) F(1,c0) F(0,c1) F(1,c1) ..  B(c0) B(c1) ..
// F(a,c): forward exchange + unpack of chunk c of array a; B(c): pack +
// inverse exchange of chunk c of the outputs.
void SlabTranspose::runSlab(fftBase *fftx, size_t A, size_t B,
                            DeviceArrays& devF, Complex **f, size_t offset,
                            double sc, const InnerSweep& inner)
{
  size_t N=std::max(A,B);
  void *st=gpu::stream();
  if(hasLocal() && !gpu::isDevice(f[0])) {
    std::cerr << "distributed convolutions need device pointers" << std::endl;
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
  size_t nc=std::max<size_t>(1,nchunks);
  if(!commStream)
    gpu::check(fftwpp_gpu_stream_create(&commStream),"stream creation");
  size_t nev=A+2*nc+2;
  while(events.size() < nev) {
    void *ev;
    gpu::check(fftwpp_gpu_event_create(&ev),"event creation");
    events.push_back(ev);
  }
  void **evX=events.data();          // A: x forward of array a done
  void **evF=events.data()+A;        // nc: chunk c of every array arrived
  void **evY=events.data()+A+nc;     // nc: y/z sweep of chunk c done
  void *evB=events[A+2*nc];          // all inverse exchanges done
  void *evStart=events[A+2*nc+1];    // previous work on the compute stream

  std::vector<Complex *> T(N);
  for(size_t a=0; a < N; ++a) T[a]=(Complex *) devT.ptr[a];

  // the exchange stream must not run ahead of earlier work on the buffers
  gpu::check(fftwpp_gpu_event_record(evStart,st),"event");
  gpu::check(fftwpp_gpu_stream_wait_event(commStream,evStart),"wait");

  for(size_t a=0; a < A; ++a) {
    if(hasLocal()) // ranks without local data have no x pass
      gpu::check(fftwpp_gpu_forward(fftx->plan(),0,nsub,1,f[a]+offset,
                                    devF.ptr[a],1,0,0,st),"forward");
    gpu::check(fftwpp_gpu_event_record(evX[a],st),"event");
  }
  for(size_t c=0; c < nc; ++c) {
    for(size_t a=0; a < A; ++a) {
      if(c == 0)
        gpu::check(fftwpp_gpu_stream_wait_event(commStream,evX[a]),"wait");
      transposeForward(devF.ptr[a],devT.ptr[a],c,nc,commStream);
    }
    gpu::check(fftwpp_gpu_event_record(evF[c],commStream),"event");
  }
  for(size_t c=0; c < nc; ++c) {
    size_t lo,hi;
    chunkRange(group.rank,c,nc,&lo,&hi);
    gpu::check(fftwpp_gpu_stream_wait_event(st,evF[c]),"wait");
    if(hi > lo) inner(T.data(),lo,hi);
    gpu::check(fftwpp_gpu_event_record(evY[c],st),"event");
    gpu::check(fftwpp_gpu_stream_wait_event(commStream,evY[c]),"wait");
    for(size_t b=0; b < B; ++b)
      transposeBackward(devT.ptr[b],devF.ptr[b],c,nc,commStream);
  }
  gpu::check(fftwpp_gpu_event_record(evB,commStream),"event");
  gpu::check(fftwpp_gpu_stream_wait_event(st,evB),"wait");
  for(size_t b=0; b < B && hasLocal(); ++b)
    gpu::check(fftwpp_gpu_backward(fftx->plan(),0,nsub,1,devF.ptr[b],
                                   f[b]+offset,0,sc,1,0,0,st),"backward");
}

void Convolution3MPI::runMPI(Complex **f, size_t offset, double sc)
{
  runSlab(fftx,A,B,devF,f,offset,sc,[this](Complex **T, size_t lo, size_t hi) {
    convolveyz[0]->convolvePlanes(T,lo*d.Y*d.Z,hi-lo,d.Y*d.Z,1.0);
  });
}

// One-time collective setup of the fused exchange: allocate the landing
// buffers, exchange their CUDA IPC handles, map the peers' buffers and build
// the per-row destination tables of both mapped passes.
void Convolution3MPI::setupFused()
{
  size_t N=std::max(A,B);
  void *st=gpu::stream();
  int P=group.size;
  size_t rows=fftx->allRows();
  size_t slabWords=rows*fftx->S;               // X x y x Z
  size_t tWords=std::max<size_t>(d.x,1)*d.Y*d.Z;
  devF.ensure(N,slabWords*sizeof(Complex));
  devT.ensure(N,tWords*sizeof(Complex));

  // all-gather the IPC handles of devT[0..N) and devF[0..B)
  size_t nh=N+B;
  std::vector<char> mine(64*nh), all(64*nh*P);
  for(size_t a=0; a < N; ++a)
    gpu::check(fftwpp_gpu_ipc_get_handle(devT.ptr[a],&mine[64*a]),"ipc handle");
  for(size_t b=0; b < B; ++b)
    gpu::check(fftwpp_gpu_ipc_get_handle(devF.ptr[b],&mine[64*(N+b)]),
               "ipc handle");
  DeviceArrays tmp;
  tmp.ensure(2,64*nh*P);
  gpu::check(fftwpp_gpu_memcpy_h2d(tmp.ptr[0],mine.data(),mine.size(),st),
             "h2d");
  gpu::check(fftwpp_gpu_comm_allgather(group.comm,tmp.ptr[0],tmp.ptr[1],64*nh,
                                       st),"all-gather (ipc handles)");
  gpu::check(fftwpp_gpu_memcpy_d2h(all.data(),tmp.ptr[1],all.size(),st),"d2h");
  gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  peerT.assign((size_t) P*N,NULL);
  peerF.assign((size_t) P*B,NULL);
  for(int p=0; p < P; ++p) {
    for(size_t k=0; k < nh; ++k) {
      void *ptr=NULL;
      if(p == group.rank)
        ptr=k < N ? devT.ptr[k] : devF.ptr[k-N];
      else {
        gpu::check(fftwpp_gpu_ipc_open(&all[64*(nh*p+k)],&ptr),"ipc open");
        opened.push_back(ptr);
      }
      if(k < N) peerT[(size_t) p*N+k]=ptr;
      else peerF[(size_t) p*B+(k-N)]=ptr;
    }
  }

  // row maps: [A forward maps of X rows][B backward maps of Y rows], each a
  // base array (uint64) followed by a stride array (int64)
  size_t fwdLen=d.X, bwdLen=d.Y;
  size_t words=A*2*fwdLen+B*2*bwdLen;
  std::vector<uint64_t> host(words);
  size_t pos=0;
  const uint64_t w=sizeof(Complex);
  for(size_t a=0; a < A; ++a) {
    // x forward: all-layout output row l belongs to the rank owning x row l;
    // it lands at T_p[(l-x0_p)][y0_me..][:]
    for(size_t l=0; l < fwdLen; ++l) {
      int p=(int) std::min<size_t>(l/ceilquotient(d.X,P),P-1);
      size_t px0;
      localdimension(d.X,p,P,&px0);
      host[pos+l]=(uint64_t) peerT[(size_t) p*N+a]+
        ((l-px0)*d.Y*d.Z+d.y0*d.Z)*w;
      host[pos+fwdLen+l]=0;
    }
    pos += 2*fwdLen;
  }
  for(size_t b=0; b < B; ++b) {
    // y backward: output row j of local plane i (x row x0_me+i) belongs to the
    // rank owning y row j; it lands at Fx_p[(x0_me+i)][j-y0_p][:]
    for(size_t j=0; j < bwdLen; ++j) {
      int p=(int) std::min<size_t>(j/ceilquotient(d.Y,P),P-1);
      size_t py0;
      size_t py=localdimension(d.Y,p,P,&py0);
      host[pos+j]=(uint64_t) peerF[(size_t) p*B+b]+
        (d.x0*py*d.Z+(j-py0)*d.Z)*w;
      host[pos+bwdLen+j]=(uint64_t) (int64_t) (py*d.Z);
    }
    pos += 2*bwdLen;
  }
  // the same destinations as per-peer row ranges for the TMA-staged kernels
  // (fftwpp_gpu_forward_dests / backward_dests): one tensor map per peer
  fwdDests.assign(A,std::vector<fftwpp_gpu_dest>());
  for(size_t a=0; a < A; ++a)
    for(int p=0; p < P; ++p) {
      size_t px0;
      size_t px=localdimension(d.X,p,P,&px0);
      fftwpp_gpu_dest t;
      t.base=(Complex *) peerT[(size_t) p*N+a]+d.y0*d.Z;
      t.row0=px0;
      t.rows=px;
      t.row_stride=d.Y*d.Z;
      t.plane_stride=0;
      fwdDests[a].push_back(t);
    }
  convolveyz[0]->outDests.assign(B,std::vector<fftwpp_gpu_dest>());
  for(size_t b=0; b < B; ++b)
    for(int p=0; p < P; ++p) {
      size_t py0;
      size_t py=localdimension(d.Y,p,P,&py0);
      fftwpp_gpu_dest t;
      t.base=(Complex *) peerF[(size_t) p*B+b]+d.x0*py*d.Z;
      t.row0=py0;
      t.rows=py;
      t.row_stride=d.Z;
      t.plane_stride=py*d.Z;
      convolveyz[0]->outDests[b].push_back(t);
    }
  devMap.ensure(1,words*sizeof(uint64_t));
  gpu::check(fftwpp_gpu_memcpy_h2d(devMap.ptr[0],host.data(),
                                   words*sizeof(uint64_t),st),"h2d");
  gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  const uint64_t *base=(const uint64_t *) devMap.ptr[0];
  convolveyz[0]->outBase.assign(B,NULL);
  convolveyz[0]->outStride.assign(B,NULL);
  for(size_t b=0; b < B; ++b) {
    const uint64_t *m=base+A*2*fwdLen+b*2*bwdLen;
    convolveyz[0]->outBase[b]=m;
    convolveyz[0]->outStride[b]=(const int64_t *) (m+bwdLen);
  }
  gpu::check(fftwpp_gpu_comm_barrier(group.comm,st),"barrier");
  fusedReady=true;
}

// x forward (stores into the peers' transposed buffers) | barrier |
// y forward, z fused convolution, y backward (stores into the peers' x slabs)
// | barrier | x backward.
void Convolution3MPI::runFused(Complex **f, size_t offset, double sc)
{
  size_t N=std::max(A,B);
  void *st=gpu::stream();
  if(!fusedReady) setupFused();
  const std::vector<ResidueCall>& calls=fftx->calls();
  size_t nsub=calls.back().sb0+calls.back().nsb;
  const uint64_t *base=(const uint64_t *) devMap.ptr[0];
  for(size_t a=0; a < A && d.y > 0; ++a) {
    // TMA bulk stores to the owners' buffers where the plan has such a
    // kernel, else per-thread stores through the row maps
    int rc=fftwpp_gpu_forward_dests(fftx->plan(),0,nsub,f[a]+offset,
                                    fwdDests[a].data(),(int) fwdDests[a].size(),
                                    0,1,0,st);
    if(rc == FFTWPP_GPU_EUNSUPPORTED) {
      const uint64_t *m=base+a*2*d.X;
      rc=fftwpp_gpu_forward_mapped(fftx->plan(),0,nsub,f[a]+offset,m,
                                   (const int64_t *) (m+d.X),1,0,st);
    }
    gpu::check(rc,"forward (fused exchange)");
  }
  gpu::check(fftwpp_gpu_comm_barrier(group.comm,st),"barrier");
  std::vector<Complex *> T(N);
  for(size_t a=0; a < N; ++a) T[a]=(Complex *) devT.ptr[a];
  if(d.x > 0)
    convolveyz[0]->convolvePlanes(T.data(),0,d.x,d.Y*d.Z,1.0);
  gpu::check(fftwpp_gpu_comm_barrier(group.comm,st),"barrier");
  for(size_t b=0; b < B && d.y > 0; ++b)
    gpu::check(fftwpp_gpu_backward(fftx->plan(),0,nsub,1,devF.ptr[b],
                                   f[b]+offset,0,sc,1,0,0,st),"backward");
}

void Convolution3MPI::HermitianSymmetrizeXY(Complex *f)
{
  if(d.y > 0 && !gpu::isDevice(f)) {
    std::cerr << "distributed convolutions need device pointers" << std::endl;
    exit(-1);
  }
  void *st=gpu::stream();
  const int P=group.size;
  const size_t Ly=d.Y, Z=d.Z, dy=d.y;
  const size_t ymax=ceilquotient(Ly,(size_t) P);
  const size_t n=Lx*ymax;
  const size_t w=sizeof(Complex);
  // z=0 entries of the local slab, padded to ymax rows of y per x
  std::vector<Complex> mine(n,Complex(0.0,0.0)), all(n*P);
  for(size_t i=0; i < Lx && dy > 0; ++i)
    gpu::check(fftwpp_gpu_memcpy2d(&mine[i*ymax],w,f+i*dy*Z,Z*w,w,dy,1,st),
               "d2h (z=0 plane)");
  DeviceArrays tmp;
  tmp.ensure(1,n*w*(P+1));
  char *send=(char *) tmp.ptr[0];
  char *recv=send+n*w;
  gpu::check(fftwpp_gpu_memcpy_h2d(send,mine.data(),n*w,st),"h2d");
  gpu::check(fftwpp_gpu_comm_allgather(group.comm,send,recv,n*w,st),
             "all-gather (z=0 plane)");
  gpu::check(fftwpp_gpu_memcpy_d2h(all.data(),recv,n*w*P,st),"d2h");
  gpu::check(fftwpp_gpu_stream_sync(st),"sync");
  std::vector<Complex> plane(Lx*Ly);
  for(int p=0; p < P; ++p) {
    size_t py0;
    size_t py=localdimension(Ly,p,P,&py0);
    for(size_t i=0; i < Lx; ++i)
      for(size_t j=0; j < py; ++j)
        plane[i*Ly+py0+j]=all[(size_t) p*n+i*ymax+j];
  }
  const size_t Hx=ceilquotient(Lx,2), Hy=ceilquotient(Ly,2);
  fftwpp::HermitianSymmetrizeXY(Hx,Hy,1,Lx/2,Ly/2,plane.data(),Ly,1);
  for(size_t i=0; i < Lx && dy > 0; ++i) {
    for(size_t j=0; j < dy; ++j) mine[i*ymax+j]=plane[i*Ly+d.y0+j];
    gpu::check(fftwpp_gpu_memcpy2d(f+i*dy*Z,Z*w,&mine[i*ymax],w,w,dy,0,st),
               "h2d (z=0 plane)");
  }
  // even lengths carry an unpaired Nyquist row at index 0: zero it for every z
  if(Lx/2 == Hx && dy > 0)
    gpu::check(fftwpp_gpu_memset(f,0,dy*Z*w,st),"memset");
  if(Ly/2 == Hy && d.y0 == 0 && dy > 0)
    for(size_t i=0; i < Lx; ++i)
      gpu::check(fftwpp_gpu_memset(f+i*dy*Z,0,Z*w,st),"memset");
  gpu::check(fftwpp_gpu_stream_sync(st),"sync");
}

void Convolution3MPI::convolveRaw(Complex **f, size_t offset, Indices *)
{
  if(inner) runPencil(f,offset,1.0);
  else if(fused) runFused(f,offset,1.0);
  else runMPI(f,offset,1.0);
}

void Convolution3MPI::convolve(Complex **f, size_t offset)
{
  if(inner) runPencil(f,offset,scale);
  else if(fused) runFused(f,offset,scale);
  else runMPI(f,offset,scale);
}

} // namespace fftwpp
