// comm.cu -- NCCL all-to-all for the distributed transpose, behind the thin C
// ABI.  Replaces the reference's adaptive MPI transpose
// (mpi/mpitranspose.h:132-161,632-931: MPI_Ialltoall or scheduled
// Isend/Irecv, optional two-stage a x b sub-blocking tuned by a latency probe).
// On one NVSwitch box every peer is at full bandwidth, so the exchange is a
// single grouped ncclSend/ncclRecv with per-peer counts (uneven splits,
// mpitranspose.h:118-130); no latency tuning, no sub-blocking.
//
// NCCL is resolved with dlopen at first use so that single-GPU users of
// lib_fftwpp.so need no NCCL at all and, under torch, the library torch
// already loaded (same soname) is reused.
#include "gpu_internal.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace fftwpp_gpu {

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct {char internal[128];} ncclUniqueId;
typedef int ncclResult_t;
const int ncclChar=0;

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t,
                            cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t,
                            cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char *(*GetErrorString)(ncclResult_t);
  bool ok;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void loadNccl()
{
  memset(&g_nccl,0,sizeof(g_nccl));
  const char *names[]={"libnccl.so.2","libnccl.so",NULL};
  void *h=NULL;
  for(int i=0; names[i] && !h; ++i)
    h=dlopen(names[i],RTLD_NOW | RTLD_GLOBAL);
  if(!h) return;
#define SYM(field, name) \
  *(void **) (&g_nccl.field)=dlsym(h,name); \
  if(!g_nccl.field) return;
  SYM(GetUniqueId,"ncclGetUniqueId")
  SYM(CommInitRank,"ncclCommInitRank")
  SYM(CommDestroy,"ncclCommDestroy")
  SYM(Send,"ncclSend")
  SYM(Recv,"ncclRecv")
  SYM(AllReduce,"ncclAllReduce")
  SYM(AllGather,"ncclAllGather")
  SYM(GroupStart,"ncclGroupStart")
  SYM(GroupEnd,"ncclGroupEnd")
  SYM(GetErrorString,"ncclGetErrorString")
#undef SYM
  g_nccl.ok=true;
}

int needNccl()
{
  std::call_once(g_nccl_once,loadNccl);
  if(!g_nccl.ok) {
    set_error("NCCL (libnccl.so.2) could not be loaded: %s",dlerror());
    return FFTWPP_GPU_ENCCL;
  }
  return 0;
}

int ncclFail(ncclResult_t r, const char *what)
{
  set_error("%s: %s",what,g_nccl.GetErrorString(r));
  return FFTWPP_GPU_ENCCL;
}

struct Comm {
  ncclComm_t comm;
  int rank, size;
  int *flag; // device scratch of the stream barrier
};

} // namespace

} // namespace fftwpp_gpu

using namespace fftwpp_gpu;

extern "C" {

int fftwpp_gpu_comm_unique_id(char *id128)
{
  int rc=needNccl();
  if(rc) return rc;
  ncclUniqueId id;
  ncclResult_t r=g_nccl.GetUniqueId(&id);
  if(r) return ncclFail(r,"ncclGetUniqueId");
  memcpy(id128,id.internal,128);
  return 0;
}

int fftwpp_gpu_comm_create(int rank, int size, const char *id128, void **comm)
{
  int rc=needNccl();
  if(rc) return rc;
  if(!comm || !id128 || rank < 0 || rank >= size) return FFTWPP_GPU_EINVAL;
  ncclUniqueId id;
  memcpy(id.internal,id128,128);
  Comm *c=new Comm;
  c->rank=rank;
  c->size=size;
  ncclResult_t r=g_nccl.CommInitRank(&c->comm,size,id,rank);
  if(r) {
    delete c;
    return ncclFail(r,"ncclCommInitRank");
  }
  c->flag=NULL;
  cudaError_t e=cudaMalloc((void **) &c->flag,2*sizeof(int));
  if(e != cudaSuccess) {
    g_nccl.CommDestroy(c->comm);
    delete c;
    return cuda_fail(e,"cudaMalloc(barrier flag)");
  }
  cudaMemset(c->flag,0,2*sizeof(int));
  *comm=c;
  return 0;
}

int fftwpp_gpu_comm_destroy(void *comm)
{
  Comm *c=(Comm *) comm;
  if(!c) return 0;
  if(g_nccl.ok) g_nccl.CommDestroy(c->comm);
  if(c->flag) cudaFree(c->flag);
  delete c;
  return 0;
}

// Stream-ordered barrier: work enqueued after it on `stream` starts only
// after every rank's work enqueued before its own barrier has completed.
int fftwpp_gpu_comm_barrier(void *comm, void *stream)
{
  int rc=needNccl();
  if(rc) return rc;
  Comm *c=(Comm *) comm;
  if(!c) return FFTWPP_GPU_EINVAL;
  const int ncclInt=2, ncclSum=0;
  ncclResult_t r=g_nccl.AllReduce(c->flag,c->flag+1,1,ncclInt,ncclSum,c->comm,
                                  (cudaStream_t) stream);
  if(r) return ncclFail(r,"ncclAllReduce(barrier)");
  return 0;
}

// recv (size*bytes, device) = concatenation of every rank's send (bytes).
int fftwpp_gpu_comm_allgather(void *comm, const void *send, void *recv,
                              uint64_t bytes, void *stream)
{
  int rc=needNccl();
  if(rc) return rc;
  Comm *c=(Comm *) comm;
  if(!c) return FFTWPP_GPU_EINVAL;
  ncclResult_t r=g_nccl.AllGather(send,recv,bytes,ncclChar,c->comm,
                                  (cudaStream_t) stream);
  if(r) return ncclFail(r,"ncclAllGather");
  return 0;
}

int fftwpp_gpu_comm_rank(void *comm) {return comm ? ((Comm *) comm)->rank : -1;}
int fftwpp_gpu_comm_size(void *comm) {return comm ? ((Comm *) comm)->size : 0;}

// All-to-all with per-peer byte counts and displacements (an MPI_Alltoallv):
// the block for peer p is send+sdispl[p] (scount[p] bytes); the block from
// peer p lands at recv+rdispl[p] (rcount[p] bytes).  The self block is a
// device-to-device copy.
int fftwpp_gpu_comm_alltoallv(void *comm, const void *send,
                              const uint64_t *scount, const uint64_t *sdispl,
                              void *recv, const uint64_t *rcount,
                              const uint64_t *rdispl, void *stream)
{
  int rc=needNccl();
  if(rc) return rc;
  Comm *c=(Comm *) comm;
  if(!c) return FFTWPP_GPU_EINVAL;
  cudaStream_t st=(cudaStream_t) stream;
  int me=c->rank;
  if(scount[me] != rcount[me]) {
    set_error("alltoallv: self send/recv counts differ");
    return FFTWPP_GPU_EINVAL;
  }
  if(scount[me]) {
    cudaError_t e=cudaMemcpyAsync((char *) recv+rdispl[me],
                                  (const char *) send+sdispl[me],scount[me],
                                  cudaMemcpyDeviceToDevice,st);
    if(e != cudaSuccess) return cuda_fail(e,"cudaMemcpyAsync(self block)");
  }
  ncclResult_t r=g_nccl.GroupStart();
  if(r) return ncclFail(r,"ncclGroupStart");
  // a failing send/recv must not leave the NCCL group open for later calls
  ncclResult_t bad=(ncclResult_t) 0;
  const char *what=NULL;
  for(int p=0; p < c->size && !bad; ++p) {
    if(p == me) continue;
    if(scount[p]) {
      r=g_nccl.Send((const char *) send+sdispl[p],scount[p],ncclChar,p,c->comm,
                    st);
      if(r) {bad=r; what="ncclSend"; break;}
    }
    if(rcount[p]) {
      r=g_nccl.Recv((char *) recv+rdispl[p],rcount[p],ncclChar,p,c->comm,st);
      if(r) {bad=r; what="ncclRecv"; break;}
    }
  }
  r=g_nccl.GroupEnd();
  if(bad) return ncclFail(bad,what);
  if(r) return ncclFail(r,"ncclGroupEnd");
  return 0;
}

}
