// capi.cu -- extern "C" entry points declared in include/fftwpp_gpu.h.
#include "gpu_internal.h"

#include <cstring>

using namespace fftwpp_gpu;

#define CUDA_TRY(call, what)                       \
  do {                                             \
    cudaError_t e_=(call);                         \
    if(e_ != cudaSuccess) return cuda_fail(e_,what); \
  } while(0)

extern "C" {

int fftwpp_gpu_device_count(void)
{
  int n=0;
  cudaError_t e=cudaGetDeviceCount(&n);
  if(e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int fftwpp_gpu_set_device(int device)
{
  CUDA_TRY(cudaSetDevice(device),"cudaSetDevice");
  return 0;
}

int fftwpp_gpu_malloc(void **ptr, size_t bytes)
{
  if(!ptr) return FFTWPP_GPU_EINVAL;
  *ptr=NULL;
  if(bytes == 0) return 0;
  CUDA_TRY(cudaMalloc(ptr,bytes),"cudaMalloc");
  return 0;
}

int fftwpp_gpu_free(void *ptr)
{
  if(!ptr) return 0;
  CUDA_TRY(cudaFree(ptr),"cudaFree");
  return 0;
}

int fftwpp_gpu_malloc_host(void **ptr, size_t bytes)
{
  if(!ptr) return FFTWPP_GPU_EINVAL;
  *ptr=NULL;
  if(bytes == 0) return 0;
  CUDA_TRY(cudaMallocHost(ptr,bytes),"cudaMallocHost");
  return 0;
}

int fftwpp_gpu_free_host(void *ptr)
{
  if(!ptr) return 0;
  CUDA_TRY(cudaFreeHost(ptr),"cudaFreeHost");
  return 0;
}

int fftwpp_gpu_memcpy_h2d(void *dst, const void *src, size_t bytes,
                          void *stream)
{
  if(bytes == 0) return 0;
  CUDA_TRY(cudaMemcpyAsync(dst,src,bytes,cudaMemcpyHostToDevice,
                           (cudaStream_t) stream),"cudaMemcpyAsync(h2d)");
  return 0;
}

int fftwpp_gpu_memcpy_d2h(void *dst, const void *src, size_t bytes,
                          void *stream)
{
  if(bytes == 0) return 0;
  CUDA_TRY(cudaMemcpyAsync(dst,src,bytes,cudaMemcpyDeviceToHost,
                           (cudaStream_t) stream),"cudaMemcpyAsync(d2h)");
  return 0;
}

int fftwpp_gpu_memcpy_d2d(void *dst, const void *src, size_t bytes,
                          void *stream)
{
  if(bytes == 0) return 0;
  CUDA_TRY(cudaMemcpyAsync(dst,src,bytes,cudaMemcpyDeviceToDevice,
                           (cudaStream_t) stream),"cudaMemcpyAsync(d2d)");
  return 0;
}

int fftwpp_gpu_memset(void *dst, int value, size_t bytes, void *stream)
{
  if(bytes == 0) return 0;
  CUDA_TRY(cudaMemsetAsync(dst,value,bytes,(cudaStream_t) stream),
           "cudaMemsetAsync");
  return 0;
}

int fftwpp_gpu_memcpy2d(void *dst, size_t dpitch, const void *src,
                        size_t spitch, size_t width, size_t height, int kind,
                        void *stream)
{
  if(width == 0 || height == 0) return 0;
  cudaMemcpyKind k=kind == 0 ? cudaMemcpyHostToDevice :
    kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  CUDA_TRY(cudaMemcpy2DAsync(dst,dpitch,src,spitch,width,height,k,
                             (cudaStream_t) stream),"cudaMemcpy2DAsync");
  return 0;
}

int fftwpp_gpu_stream_sync(void *stream)
{
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t) stream),
           "cudaStreamSynchronize");
  return 0;
}

int fftwpp_gpu_stream_create(void **stream)
{
  cudaStream_t st;
  // highest priority: CTAs of the (small) exchange kernels are placed before
  // queued CTAs of the compute kernels whenever SM resources free up
  int least=0, greatest=0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least,&greatest),
           "cudaDeviceGetStreamPriorityRange");
  CUDA_TRY(cudaStreamCreateWithPriority(&st,cudaStreamNonBlocking,greatest),
           "cudaStreamCreateWithPriority");
  *stream=(void *) st;
  return 0;
}

int fftwpp_gpu_stream_destroy(void *stream)
{
  if(stream) CUDA_TRY(cudaStreamDestroy((cudaStream_t) stream),
                      "cudaStreamDestroy");
  return 0;
}

int fftwpp_gpu_event_create(void **event)
{
  cudaEvent_t ev;
  CUDA_TRY(cudaEventCreateWithFlags(&ev,cudaEventDisableTiming),
           "cudaEventCreate");
  *event=(void *) ev;
  return 0;
}

int fftwpp_gpu_event_destroy(void *event)
{
  if(event) CUDA_TRY(cudaEventDestroy((cudaEvent_t) event),"cudaEventDestroy");
  return 0;
}

int fftwpp_gpu_event_record(void *event, void *stream)
{
  CUDA_TRY(cudaEventRecord((cudaEvent_t) event,(cudaStream_t) stream),
           "cudaEventRecord");
  return 0;
}

int fftwpp_gpu_event_sync(void *event)
{
  CUDA_TRY(cudaEventSynchronize((cudaEvent_t) event),"cudaEventSynchronize");
  return 0;
}

int fftwpp_gpu_stream_wait_event(void *stream, void *event)
{
  CUDA_TRY(cudaStreamWaitEvent((cudaStream_t) stream,(cudaEvent_t) event,0),
           "cudaStreamWaitEvent");
  return 0;
}

int fftwpp_gpu_device_sync(void)
{
  CUDA_TRY(cudaDeviceSynchronize(),"cudaDeviceSynchronize");
  return 0;
}

int fftwpp_gpu_is_device_ptr(const void *ptr)
{
  if(!ptr) return 0;
  cudaPointerAttributes attr;
  cudaError_t e=cudaPointerGetAttributes(&attr,ptr);
  if(e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return attr.type == cudaMemoryTypeDevice ||
    attr.type == cudaMemoryTypeManaged;
}

const char *fftwpp_gpu_last_error(void)
{
  return g_err;
}

uint64_t fftwpp_gpu_launch_count(void)
{
  return g_launches.load();
}

int fftwpp_gpu_plan_create(const fftwpp_gpu_pad_desc *desc,
                           fftwpp_gpu_plan **plan)
{
  Plan *pl=NULL;
  int rc=plan_build(desc,&pl);
  if(rc) return rc;
  *plan=(fftwpp_gpu_plan *) pl;
  return 0;
}

int fftwpp_gpu_plan_set_tag(fftwpp_gpu_plan *plan, int tag)
{
  if(!plan || tag < 0 || tag >= PROF_KEYS/4) return FFTWPP_GPU_EINVAL;
  ((Plan *) plan)->tag=tag;
  return 0;
}

int fftwpp_gpu_plan_set_outer(fftwpp_gpu_plan *child, fftwpp_gpu_plan *parent,
                              uint64_t n)
{
  Plan *c=(Plan *) child;
  Plan *p=(Plan *) parent;
  if(!c || !p || n == 0) return FFTWPP_GPU_EINVAL;
  if(!c->fast || c->dev.C < 2) {
    set_error("set_outer: the child pass must be a power-of-two strided pass");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  c->dev.oz1=p->dev.z1;
  c->dev.oz2=p->dev.z2;
  c->dev.ozshift=p->dev.zshift;
  c->dev.on=(long long) n;
  c->dev.oen=1;
  return 0;
}

int fftwpp_gpu_profile_enable(int on) {return prof_enable(on);}

int fftwpp_gpu_profile_read(double *ms, uint64_t *count)
{
  if(!ms || !count) return FFTWPP_GPU_EINVAL;
  return prof_read(ms,count);
}

int fftwpp_gpu_plan_destroy(fftwpp_gpu_plan *plan)
{
  Plan *pl=(Plan *) plan;
  if(pl) {
    fast_plan_free(pl);
    delete pl;
  }
  return 0;
}

static int check_range(Plan *pl, uint64_t sb0, uint64_t nsb)
{
  if(!pl || nsb == 0 || sb0+nsb > pl->hsub.size()) {
    set_error("invalid sub-block range [%llu,+%llu)",
              (unsigned long long) sb0,(unsigned long long) nsb);
    return FFTWPP_GPU_EINVAL;
  }
  return 0;
}

int fftwpp_gpu_forward(fftwpp_gpu_plan *plan, uint64_t sb0, uint64_t nsb,
                       int all_layout, const void *f, void *F, uint64_t nrows,
                       uint64_t f_rowstride, uint64_t F_rowstride,
                       void *stream)
{
  Plan *pl=(Plan *) plan;
  int rc=check_range(pl,sb0,nsb);
  if(rc) return rc;
  rc=fast_try_forward(pl,sb0,nsb,all_layout,f,F,nrows,f_rowstride,
                      F_rowstride,(cudaStream_t) stream);
  if(rc != 0) return rc < 0 ? rc : 0;
  if(pl->dev.oen) {
    set_error("forward: outer-twiddle plans need the power-of-two fast path");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  return generic_forward(pl,sb0,nsb,all_layout,f,F,nrows,f_rowstride,
                         F_rowstride,(cudaStream_t) stream);
}

int fftwpp_gpu_backward(fftwpp_gpu_plan *plan, uint64_t sb0, uint64_t nsb,
                        int all_layout, const void *F, void *f,
                        int accumulate, double scale, uint64_t nrows,
                        uint64_t F_rowstride, uint64_t f_rowstride,
                        void *stream)
{
  Plan *pl=(Plan *) plan;
  int rc=check_range(pl,sb0,nsb);
  if(rc) return rc;
  rc=fast_try_backward(pl,sb0,nsb,all_layout,F,f,accumulate,scale,nrows,
                       F_rowstride,f_rowstride,(cudaStream_t) stream);
  if(rc != 0) return rc < 0 ? rc : 0;
  if(pl->dev.oen) {
    set_error("backward: outer-twiddle plans need the power-of-two fast path");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  return generic_backward(pl,sb0,nsb,all_layout,F,f,accumulate,scale,nrows,
                          F_rowstride,f_rowstride,(cudaStream_t) stream);
}

int fftwpp_gpu_forward_mapped(fftwpp_gpu_plan *plan, uint64_t sb0,
                              uint64_t nsb, const void *f,
                              const uint64_t *rowbase,
                              const int64_t *rowstride, uint64_t nrows,
                              uint64_t f_rowstride, void *stream)
{
  Plan *pl=(Plan *) plan;
  int rc=check_range(pl,sb0,nsb);
  if(rc) return rc;
  if(!rowbase || !rowstride) return FFTWPP_GPU_EINVAL;
  rc=fast_try_forward(pl,sb0,nsb,1,f,NULL,nrows,f_rowstride,0,
                      (cudaStream_t) stream,
                      (const unsigned long long *) rowbase,
                      (const long long *) rowstride);
  if(rc == 0) {
    set_error("forward_mapped: needs the power-of-two strided fast path");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  return rc < 0 ? rc : 0;
}

int fftwpp_gpu_backward_mapped(fftwpp_gpu_plan *plan, uint64_t sb0,
                               uint64_t nsb, const void *F,
                               const uint64_t *rowbase,
                               const int64_t *rowstride, uint64_t plane0,
                               double scale, uint64_t nrows,
                               uint64_t F_rowstride, void *stream)
{
  Plan *pl=(Plan *) plan;
  int rc=check_range(pl,sb0,nsb);
  if(rc) return rc;
  if(!rowbase || !rowstride) return FFTWPP_GPU_EINVAL;
  rc=fast_try_backward(pl,sb0,nsb,1,F,NULL,0,scale,nrows,F_rowstride,0,
                       (cudaStream_t) stream,
                       (const unsigned long long *) rowbase,
                       (const long long *) rowstride,(long long) plane0);
  if(rc == 0) {
    set_error("backward_mapped: needs the uniform complex power-of-two fast "
              "path");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  return rc < 0 ? rc : 0;
}

int fftwpp_gpu_forward_dests(fftwpp_gpu_plan *plan, uint64_t sb0, uint64_t nsb,
                             const void *f, const fftwpp_gpu_dest *dests,
                             int ndest, uint64_t plane0, uint64_t nrows,
                             uint64_t f_rowstride, void *stream)
{
  Plan *pl=(Plan *) plan;
  int rc=check_range(pl,sb0,nsb);
  if(rc) return rc;
  if(!dests || ndest < 1) return FFTWPP_GPU_EINVAL;
  rc=tma_try_forward(pl,sb0,nsb,1,f,NULL,nrows,f_rowstride,0,
                     (cudaStream_t) stream,dests,ndest,plane0);
  if(rc == 0) {
    set_error("forward_dests: the plan has no TMA-staged forward kernel");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  return rc < 0 ? rc : 0;
}

int fftwpp_gpu_backward_dests(fftwpp_gpu_plan *plan, uint64_t sb0,
                              uint64_t nsb, const void *F,
                              const fftwpp_gpu_dest *dests, int ndest,
                              uint64_t plane0, double scale, uint64_t nrows,
                              uint64_t F_rowstride, void *stream)
{
  Plan *pl=(Plan *) plan;
  int rc=check_range(pl,sb0,nsb);
  if(rc) return rc;
  if(!dests || ndest < 1) return FFTWPP_GPU_EINVAL;
  rc=tma_try_backward(pl,sb0,nsb,1,F,NULL,0,scale,nrows,F_rowstride,0,
                      (cudaStream_t) stream,dests,ndest,plane0);
  if(rc == 0) {
    set_error("backward_dests: the plan has no TMA-staged backward kernel");
    return FFTWPP_GPU_EUNSUPPORTED;
  }
  return rc < 0 ? rc : 0;
}

int fftwpp_gpu_mapped_supported(fftwpp_gpu_plan *plan, int backward)
{
  return plan ? fast_mapped_supported((Plan *) plan,backward) : 0;
}

int fftwpp_gpu_ipc_get_handle(void *devptr, char *handle64)
{
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h,devptr),"cudaIpcGetMemHandle");
  memcpy(handle64,&h,sizeof(h));
  return 0;
}

int fftwpp_gpu_ipc_open(const char *handle64, void **peerptr)
{
  cudaIpcMemHandle_t h;
  memcpy(&h,handle64,sizeof(h));
  CUDA_TRY(cudaIpcOpenMemHandle(peerptr,h,cudaIpcMemLazyEnablePeerAccess),
           "cudaIpcOpenMemHandle");
  return 0;
}

int fftwpp_gpu_ipc_close(void *peerptr)
{
  if(peerptr) CUDA_TRY(cudaIpcCloseMemHandle(peerptr),"cudaIpcCloseMemHandle");
  return 0;
}

int fftwpp_gpu_convolve(fftwpp_gpu_plan *plan, void *const *f, uint32_t A,
                        uint32_t B, int mult, double scale, uint64_t nrows,
                        uint64_t rowstride, void *stream)
{
  Plan *pl=(Plan *) plan;
  if(!pl || !f || A == 0 || B == 0 || A > MAXARRAYS || B > MAXARRAYS) {
    set_error("convolve: invalid arguments");
    return FFTWPP_GPU_EINVAL;
  }
  if((mult == FFTWPP_MULT_BINARY || mult == FFTWPP_MULT_REALBINARY ||
      mult == FFTWPP_MULT_CORRELATION) && (A != 2 || B != 1)) {
    set_error("convolve: binary multipliers need A=2, B=1");
    return FFTWPP_GPU_EINVAL;
  }
  if(mult == FFTWPP_MULT_NONE && B > A) {
    set_error("convolve: multNone needs B <= A");
    return FFTWPP_GPU_EINVAL;
  }
  int rc=tmem_try_convolve(pl,f,A,B,mult,scale,nrows,rowstride,
                           (cudaStream_t) stream);
  if(rc != 0) return rc < 0 ? rc : 0;
  rc=fast_try_convolve(pl,f,A,B,mult,scale,nrows,rowstride,
                       (cudaStream_t) stream);
  if(rc != 0) return rc < 0 ? rc : 0;
  return generic_convolve(pl,f,A,B,mult,scale,nrows,rowstride,
                          (cudaStream_t) stream);
}

int fftwpp_gpu_scale(double *x, double scale, uint64_t n0, uint64_t n1,
                     uint64_t n2, uint64_t s0, uint64_t s1, void *stream)
{
  return launch_scale(x,scale,n0,n1,n2,s0,s1,(cudaStream_t) stream);
}

int fftwpp_gpu_copy3(void *dst, const void *src, uint64_t n0, uint64_t n1,
                     uint64_t n2, uint64_t d0, uint64_t d1, uint64_t s0,
                     uint64_t s1, void *stream)
{
  return launch_copy3(dst,src,n0,n1,n2,d0,d1,s0,s1,(cudaStream_t) stream);
}

}
