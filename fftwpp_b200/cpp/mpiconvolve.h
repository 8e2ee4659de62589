/* mpiconvolve.h -- distributed 2-D/3-D hybrid convolutions over NCCL.
 *
 * Mirrors the reference's MPI layer for the slab decomposition
 * (mpi/mpigroup.h:8-167, mpi/mpitranspose.h:118-130, mpi/mpiconvolve.h:182-305):
 * every rank holds all of x, a contiguous slice of y and all of z; the x pass
 * is local; the x-transformed rows are redistributed so that each rank owns a
 * slice of transformed x rows with all of y (the reference's localize1); the
 * y-z sub-convolutions are local; the inverse exchange (localize0) precedes
 * the local x backward pass.  The adaptive MPI transpose is replaced by one
 * NCCL grouped send/receive per array and direction.
 *
 * Pencil decomposition (reference mpi/mpiconvolve.h:208-216: a distributed
 * y-z convolution nested inside each transformed x row, over a second
 * communicator; process grid as mpi/mpigroup.h:33-50) is available when a
 * second group is given: y is split over the first group, z over the second.
 * The reference only selects it when ranks > Ly, so on one box it is a forced
 * test / bench mode.  Pencil runs use the NCCL exchanges.
 */
#ifndef FFTWPP_B200_MPICONVOLVE_H
#define FFTWPP_B200_MPICONVOLVE_H

#include <algorithm>
#include <cstdint>
#include <functional>
#include <vector>

#include "convolve.h"
#include "../../include/fftwpp_gpu.h"

namespace utils {

// Extent and start of rank's share of N items: n=ceil(N/size), start=n*rank,
// last ranks short or empty (reference mpitranspose.h:118-130).
inline size_t localdimension(size_t N, size_t rank, size_t size, size_t *start)
{
  size_t n=ceilquotient(N,size);
  size_t s=n*rank;
  if(start) *start=s < N ? s : N;
  if(s >= N) return 0;
  return s+n <= N ? n : N-s;
}

// Process group: the NCCL counterpart of the reference MPIgroup.
class MPIgroup {
public:
  int rank,size;
  void *comm; // fftwpp_gpu_comm handle
  MPIgroup(int rank, int size, void *comm) : rank(rank), size(size),
                                             comm(comm) {}
};

// Local dimensions of the slab decomposition (reference split3 with xy=true):
// input is X x y x Z, the transposed data is x x Y x Z.
class split3 {
public:
  size_t X,Y,Z;    // global: transformed x rows, y length, z length (words)
  size_t x,y,z;    // local extents
  size_t x0,y0,z0; // local offsets
  split3() {}
  split3(size_t X, size_t Y, size_t Z, const MPIgroup& group) :
    X(X), Y(Y), Z(Z), z(Z), z0(0) {
    x=localdimension(X,group.rank,group.size,&x0);
    y=localdimension(Y,group.rank,group.size,&y0);
  }
};

// 2-D counterpart (reference mpi/mpigroup.h:69-103): local matrix X x y,
// transposed x x Y, n = words of the larger of the two.
class split {
public:
  size_t X,Y;
  size_t x,y;
  size_t x0,y0;
  size_t n;
  split() {}
  split(size_t X, size_t Y, const MPIgroup& group) : X(X), Y(Y) {
    x=localdimension(X,group.rank,group.size,&x0);
    y=localdimension(Y,group.rank,group.size,&y0);
    n=std::max(X*y,x*Y);
  }
};

}

namespace fftwpp {

// Output-buffer decomposition and allocation helpers with the reference's
// names (mpi/mpiconvolve.h:36-70).  X = l*D transformed rows per residue call,
// Y / Z = stored input lengths of the inner dimensions.  The buffers these
// sizes describe are advisory here (device scratch is owned by the
// convolution objects); outputBuffer returns HOST arrays for callers that
// pass them to the constructors.
inline utils::split outputSplit(fftBase *fftx, fftBase *ffty,
                                const utils::MPIgroup& group)
{
  return utils::split(fftx->l*fftx->D,ffty->inputLength(),group);
}

inline utils::split3 outputSplit(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                                 const utils::MPIgroup& group)
{
  return utils::split3(fftx->l*fftx->D,ffty->inputLength(),
                       fftz->inputLength(),group);
}

inline size_t bufferSize(fftBase *fftx, fftBase *ffty,
                         const utils::MPIgroup& group)
{
  return outputSplit(fftx,ffty,group).n;
}

inline size_t bufferSize(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                         const utils::MPIgroup& group)
{
  utils::split3 d=outputSplit(fftx,ffty,fftz,group);
  return std::max(d.X*d.y,d.x*d.Y)*d.Z;
}

inline Complex **outputBuffer(fftBase *fftx, fftBase *ffty,
                              const utils::MPIgroup& group)
{
  return utils::ComplexAlign(std::max(fftx->app.A,fftx->app.B),
                             bufferSize(fftx,ffty,group));
}

inline Complex **outputBuffer(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                              const utils::MPIgroup& group)
{
  return utils::ComplexAlign(std::max(fftx->app.A,fftx->app.B),
                             bufferSize(fftx,ffty,fftz,group));
}

// The slab exchange shared by the distributed convolutions: global data of
// X transformed x rows by Y rows of Z words; before the exchange every rank
// holds all X rows of its y slice (X x y x Z), after it all of Y for its slice
// of x rows (x x Y x Z) -- the reference's mpitranspose localize1/localize0
// (mpi/mpitranspose.h:163-933) as one NCCL grouped send/receive per array.
class SlabTranspose {
public:
  utils::MPIgroup group;
  utils::split3 d;

  // byte counts/displacements of the two exchanges (for tests); chunk c of
  // nchunks restricts every rank's transformed x rows to its c-th sub-range
  void exchangeTable(int direction, uint64_t *scount, uint64_t *sdispl,
                     uint64_t *rcount, uint64_t *rdispl, size_t chunk=0,
                     size_t nchunks=1);

  // Number of x-row chunks the exchanges are pipelined in (overlap of the
  // all-to-all with the inner sweep, cf. mpi/mpiconvolve.h:125-139); env
  // FFTWPP_MPI_CHUNKS, default 1 (measured: no gain from chunking).
  size_t nchunks;

  // false on ranks that hold no input data (no y rows, or -- pencil -- an
  // empty z slice): they skip the local x passes and still take part in the
  // exchanges and the inner sweep
  bool hasLocal() const {return d.y > 0 && !emptyZ;}

protected:
  bool emptyZ;
  SlabTranspose(const utils::MPIgroup& group);
  ~SlabTranspose();
  DeviceArrays devT;   // transposed data: x x Y x Z per array
  DeviceArrays devP;   // pack/unpack staging
  void *commStream;
  std::vector<void *> events;
  // inner(T,lo,hi): convolve the local transformed x rows [lo,hi) of the
  // arrays T[a] (x x Y x Z each) in place
  typedef std::function<void(Complex **T, size_t lo, size_t hi)> InnerSweep;
  void runSlab(fftBase *fftx, size_t A, size_t B, DeviceArrays& devF,
               Complex **f, size_t offset, double scale,
               const InnerSweep& inner);
  void chunkRange(int rank, size_t c, size_t nc, size_t *lo, size_t *hi);
  bool agree(bool mine);
  void transposeForward(void *Fx, void *T, size_t c, size_t nc, void *st);
  void transposeBackward(void *T, void *Fx, size_t c, size_t nc, void *st);
};

// The nested exchange of the pencil decomposition, batched over the local
// transformed x rows ("planes"): before it every rank of the group holds, for
// each plane, all R rows of its z slice (planes x R x z); after it its slice
// of the rows with all of z (planes x r x Z) -- the reference's inner
// mpitranspose over communicator2 (mpi/mpiconvolve.h:72-179 inside :208-216).
class BatchedTranspose {
public:
  utils::MPIgroup group;
  size_t R,Z;      // global: rows (transformed y rows), z length
  size_t r,z;      // local extents
  size_t r0,z0;
  BatchedTranspose(const utils::MPIgroup& group, size_t R, size_t Z);
  // U: planes x R x z  ->  V: planes x r x Z   (device, Complex words)
  void forward(const void *U, void *V, size_t planes, DeviceArrays& scratch,
               void *stream);
  void backward(const void *V, void *U, size_t planes, DeviceArrays& scratch,
                void *stream);
};

// Distributed 2-D convolution (reference Convolution2MPI,
// mpi/mpiconvolve.h:72-179): every rank holds all of x and a slice of y.
class Convolution2MPI : public Convolution2, public SlabTranspose {
public:
  // fftx must be built for the LOCAL slab: C = S = d.y; ffty as in the
  // serial case (tests: mpi/tests/hybridconv2.cc).
  Convolution2MPI(fftBase *fftx, fftBase *ffty, const utils::MPIgroup& group);

  // f: device pointers to the local slabs (Lx x d.y input words each).
  void convolveRaw(Complex **f, size_t offset=0, Indices *indices=NULL);
  void convolve(Complex **f, size_t offset=0);

  size_t stridex() {return d.Y;}
  size_t blocksizex(size_t) {return d.x;}
  size_t indexBase() {return d.x0;}
  size_t inputLengthy() {return d.y;}

protected:
  void runMPI(Complex **f, size_t offset, double scale);
};

class Convolution3MPI : public Convolution3, public SlabTranspose {
public:
  // d: X = fftx->allRows(), Y = Ly, Z = Sy (row length in words)

  // fftx must be built for the LOCAL slab: C = S = d.y*Lz; ffty/fftz as in the
  // serial case (tests: mpi/tests/hybridconvr3.cc:85-102).
  Convolution3MPI(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                  const utils::MPIgroup& group);
  // Pencil decomposition: y split over `group`, z over `groupYZ`.  fftx is
  // built for the local pencil (C = S = d.y*z), ffty for the local z slice
  // (C = S = z), fftz as in the serial case.
  Convolution3MPI(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                  const utils::MPIgroup& group, const utils::MPIgroup& groupYZ);
  virtual ~Convolution3MPI();
  bool pencil() const {return inner != NULL;}

  // f: device pointers to the local slabs (Lx x d.y x Lz input words each).
  void convolveRaw(Complex **f, size_t offset=0, Indices *indices=NULL);
  void convolve(Complex **f, size_t offset=0);

  // Fused exchange: the x forward pass and the y backward pass store their
  // results straight into the owning peer's buffers over NVLink (CUDA IPC
  // mapped), so there is no send buffer, no pack/unpack pass and no NCCL
  // copy kernel; NCCL only provides two tiny stream barriers per convolution.
  // Enabled by default when the passes are on the power-of-two fast path;
  // env FFTWPP_MPI_FUSED=0 selects the NCCL all-to-all path.
  bool fused;

  // Enforce Hermitian symmetry on the distributed centred data of a Hermitian
  // run (fftx, ffty centred, fftz Hermitian): f is this rank's DEVICE slab
  // Lx x d.y x d.Z.  Counterpart of the reference's
  // HermitianSymmetrizeXY(split3&, Complex *), mpi/mpiconvolve.cc:11-142:
  // only the z=0 plane carries a constraint, so the ranks all-gather that
  // plane, apply the serial rule (convolve.h:1208-1267) and keep their slice.
  void HermitianSymmetrizeXY(Complex *f);

protected:
  BatchedTranspose *inner;     // pencil mode: the nested y-z exchange
  DeviceArrays devU,devV,devS; // pencil mode: y-transformed planes, their
                               // transposes, pack/unpack scratch
  void init(const utils::MPIgroup& group, size_t zlocal);
  void runPencil(Complex **f, size_t offset, double scale);
  bool fusedReady;
  std::vector<void *> peerT;   // [p*N+a]: peer p's transposed buffer a
  std::vector<void *> peerF;   // [p*B+b]: peer p's x-slab landing buffer b
  std::vector<void *> opened;  // IPC mappings to close
  DeviceArrays devMap;         // row maps (base, stride) per array
  std::vector<std::vector<fftwpp_gpu_dest> > fwdDests; // [a][peer]
  void setupFused();
  void runFused(Complex **f, size_t offset, double scale);
  void runMPI(Complex **f, size_t offset, double scale);
};

}

#endif
