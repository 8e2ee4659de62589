// gpu_internal.h -- shared declarations of the CUDA layer (not installed).
#ifndef FFTWPP_GPU_INTERNAL_H
#define FFTWPP_GPU_INTERNAL_H

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <vector>

#include "../../include/fftwpp_gpu.h"

namespace fftwpp_gpu {

static const int NTHREADS=256;
static const int MAXRAD=24;    // stages of a mixed-radix plan
static const int MAXPRIME=64;  // largest prime factor handled by the O(r^2) butterfly
static const int MAXARRAYS=8;  // max(A,B) of a fused convolution

// Roots of unity and radix schedule of one FFT length.
struct FftTab {
  const double2 *omega; // omega[k]=exp(2 pi i k/n)
  // power-of-two lengths: twiddles of the register radix-8 passes; pass i
  // (ls_i=lg-3(i+1) > 0) at off_i+(u-1)*2^ls_i+j = omega^{(j*u) 8^i},
  // off_i=7*sum_{k<i} 2^ls_k, so threads of a warp read consecutive entries
  const double2 *tw8;
  int n;
  int nrad;
  int rad[MAXRAD];
};

struct SubBlockDev {
  unsigned mlen;
  unsigned nout;
  unsigned flags;
  unsigned tab;
  long long k0;
  long long off_call;
  long long off_all;
};

// Everything a kernel needs about one padded FFT, passed by value.
struct PlanDev {
  int kind;
  int Lin;          // stored input words per lane
  int jmin, jmax;   // logical index range of the input
  int C;
  int zshift;       // <0: z1[e]; else z1[e>>zshift]*z2[e&mask]
  unsigned nmask;   // N-1 when N is a power of two (and < 2^31), else 0
  int small32;      // k0*|j| < 2^32 for every sub-block and input index
  long long N;
  long long S;
  const double2 *z1;
  const double2 *z2;
  FftTab tab[2];
  // Optional outer twiddle of a two-stage (inner, p > 2) transform: output
  // row l of sub-block k0, column c is multiplied by zeta_Nbig^{(on*l+k0)*c}
  // on the way out of a forward pass (conjugate on the way into a backward
  // pass); oz1/oz2/ozshift are the parent plan's zeta tables.
  const double2 *oz1;
  const double2 *oz2;
  int ozshift;
  int oen;
  long long on;
  // Optional output row map of one launch (fused exchange): output row r of
  // plane i is written at (word *) omBase[r] + i*omStride[r] + column instead
  // of the dense layout; omBase entries may be peer-GPU addresses (NVLink).
  const unsigned long long *omBase;
  const long long *omStride;
  long long omPlane0; // plane index of the launch's first row
};

struct ConvPtrs {
  void *p[MAXARRAYS];
};

// What the specialised power-of-two kernels know about a plan
// (fast_kernels.cu: fast_plan_init).
struct FastInfo {
  int log2m;        // m = 2^log2m: the longer of the (at most two) FFT lengths
  bool uniform;     // every sub-block has length m
  int nterm;        // max number of input terms folded into one W[s]
  bool pairable;    // REAL plans: every full-length sub-block is an r2c block,
                    // so two adjacent real columns share one complex FFT
};

struct Plan {
  fftwpp_gpu_pad_desc desc;
  PlanDev dev;
  std::vector<SubBlockDev> hsub;
  const SubBlockDev *dsub;
  unsigned mmax;
  int device;
  std::vector<void *> owned;
  FastInfo *fast;
  int tag; // profiling tag set by the host layer (which pass this plan serves)
  Plan() : dsub(NULL), mmax(0), device(0), fast(NULL), tag(0) {}
  ~Plan();
};

__host__ __device__ inline double wscale(double v, double s) {return v*s;}
__host__ __device__ inline double2 wscale(double2 v, double s)
{
  return make_double2(v.x*s,v.y*s);
}

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

int plan_build(const fftwpp_gpu_pad_desc *d, Plan **out);

// Optional per-launch CUDA-event timing (fftwpp_gpu_profile_*).  key =
// 4*plan tag + op, op: 0 forward, 1 backward, 2 fused convolution, 3 other.
static const int PROF_KEYS=64;
void prof_begin(int key, cudaStream_t st);
int check_launch(const char *what, cudaStream_t st);
int sm_count();
int prof_enable(int on);
int prof_read(double *ms, uint64_t *count);

int generic_forward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                    const void *f, void *F, uint64_t nrows, uint64_t frs,
                    uint64_t Frs, cudaStream_t st);
int generic_backward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                     const void *F, void *f, int accumulate, double scale,
                     uint64_t nrows, uint64_t Frs, uint64_t frs,
                     cudaStream_t st);
int generic_convolve(Plan *pl, void *const *f, uint32_t A, uint32_t B,
                     int mult, double scale, uint64_t nrows, uint64_t rs,
                     cudaStream_t st);
int launch_scale(double *x, double scale, uint64_t n0, uint64_t n1,
                 uint64_t n2, uint64_t s0, uint64_t s1, cudaStream_t st);
int launch_copy3(void *dst, const void *src, uint64_t n0, uint64_t n1,
                 uint64_t n2, uint64_t d0, uint64_t d1, uint64_t s0,
                 uint64_t s1, cudaStream_t st);

// Specialised power-of-two kernels (fast_kernels.cu).  Each try_* returns
// 1 if it handled the request, 0 if the generic kernel must be used, <0 on
// error.
void fast_plan_init(Plan *pl);
void fast_plan_free(Plan *pl);
int fast_try_forward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                     const void *f, void *F, uint64_t nrows, uint64_t frs,
                     uint64_t Frs, cudaStream_t st,
                     const unsigned long long *omBase=NULL,
                     const long long *omStride=NULL, long long omPlane0=0);
int fast_try_backward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                      const void *F, void *f, int accumulate, double scale,
                      uint64_t nrows, uint64_t Frs, uint64_t frs,
                      cudaStream_t st,
                      const unsigned long long *omBase=NULL,
                      const long long *omStride=NULL, long long omPlane0=0);
int fast_try_convolve(Plan *pl, void *const *f, uint32_t A, uint32_t B,
                      int mult, double scale, uint64_t nrows, uint64_t rs,
                      cudaStream_t st);
int fast_mapped_supported(Plan *pl, int backward);
// fused rows with TMEM-parked spectra (tmem_kernels.cu)
int tmem_try_convolve(Plan *pl, void *const *f, uint32_t A, uint32_t B,
                      int mult, double scale, uint64_t nrows, uint64_t rs,
                      cudaStream_t st);

// TMA-staged strided passes (tma_kernels.cu): tiles move between HBM and
// shared memory with cp.async.bulk.tensor (tensor maps built per launch),
// completion on mbarriers.  Same return convention as fast_try_*.
// dests != NULL: the output rows go to `ndest` destinations (fused exchange,
// see fftwpp_gpu_forward_dests) instead of F / f.
int tma_try_forward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                    const void *f, void *F, uint64_t nrows, uint64_t frs,
                    uint64_t Frs, cudaStream_t st,
                    const fftwpp_gpu_dest *dests=NULL, int ndest=0,
                    uint64_t plane0=0);
int tma_try_backward(Plan *pl, uint64_t sb0, uint64_t nsb, int layout,
                     const void *F, void *f, int accumulate, double scale,
                     uint64_t nrows, uint64_t Frs, uint64_t frs,
                     cudaStream_t st, const fftwpp_gpu_dest *dests=NULL,
                     int ndest=0, uint64_t plane0=0);

} // namespace fftwpp_gpu

#endif
