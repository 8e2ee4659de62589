/* mpiconvolve.h -- distributed 3-D hybrid convolution over NCCL.
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
 * Pencil decomposition (Convolution2MPI inside each x row) is only selected by
 * the reference when ranks > Ly (mpigroup.h:33-39); it is not implemented here.
 */
#ifndef FFTWPP_B200_MPICONVOLVE_H
#define FFTWPP_B200_MPICONVOLVE_H

#include <cstdint>
#include <vector>

#include "convolve.h"

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

}

namespace fftwpp {

class Convolution3MPI : public Convolution3 {
public:
  utils::MPIgroup group;
  utils::split3 d;   // X = fftx->allRows(), Y = Ly, Z = Sy (row length in words)

  // fftx must be built for the LOCAL slab: C = S = d.y*Lz; ffty/fftz as in the
  // serial case (tests: mpi/tests/hybridconvr3.cc:85-102).
  Convolution3MPI(fftBase *fftx, fftBase *ffty, fftBase *fftz,
                  const utils::MPIgroup& group);
  virtual ~Convolution3MPI();

  // f: device pointers to the local slabs (Lx x d.y x Lz input words each).
  void convolveRaw(Complex **f, size_t offset=0, Indices *indices=NULL);
  void convolve(Complex **f, size_t offset=0);

  // byte counts/displacements of the two exchanges (for tests); chunk c of
  // nchunks restricts every rank's transformed x rows to its c-th sub-range
  void exchangeTable(int direction, uint64_t *scount, uint64_t *sdispl,
                     uint64_t *rcount, uint64_t *rdispl, size_t chunk=0,
                     size_t nchunks=1);

  // Number of x-row chunks the exchanges are pipelined in (overlap of the
  // all-to-all with the y/z sweep, cf. mpi/mpiconvolve.h:125-139); env
  // FFTWPP_MPI_CHUNKS, default 4.
  size_t nchunks;

  // Fused exchange: the x forward pass and the y backward pass store their
  // results straight into the owning peer's buffers over NVLink (CUDA IPC
  // mapped), so there is no send buffer, no pack/unpack pass and no NCCL
  // copy kernel; NCCL only provides two tiny stream barriers per convolution.
  // Enabled by default when the passes are on the power-of-two fast path;
  // env FFTWPP_MPI_FUSED=0 selects the NCCL all-to-all path.
  bool fused;

protected:
  bool fusedReady;
  std::vector<void *> peerT;   // [p*N+a]: peer p's transposed buffer a
  std::vector<void *> peerF;   // [p*B+b]: peer p's x-slab landing buffer b
  std::vector<void *> opened;  // IPC mappings to close
  DeviceArrays devMap;         // row maps (base, stride) per array
  void setupFused();
  void runFused(Complex **f, size_t offset, double scale);
  DeviceArrays devT;   // transposed data: x x Y x Z per array
  DeviceArrays devP;   // pack/unpack staging
  void *commStream;
  std::vector<void *> events;
  void runMPI(Complex **f, size_t offset, double scale);
  void chunkRange(int rank, size_t c, size_t nc, size_t *lo, size_t *hi);
  void transposeForward(void *Fx, void *T, size_t c, size_t nc, void *st);
  void transposeBackward(void *T, void *Fx, size_t c, size_t nc, void *st);
};

}

#endif
