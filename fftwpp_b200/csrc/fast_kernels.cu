// fast_kernels.cu -- specialised power-of-two register kernels (hot shapes).
#include "gpu_internal.h"

namespace fftwpp_gpu {

struct FastInfo {
  int dummy;
};

void fast_plan_init(Plan *pl) {pl->fast=NULL;}
void fast_plan_free(Plan *pl) {delete pl->fast; pl->fast=NULL;}

int fast_try_forward(Plan *, uint64_t, uint64_t, int, const void *, void *,
                     uint64_t, uint64_t, uint64_t, cudaStream_t)
{
  return 0;
}

int fast_try_backward(Plan *, uint64_t, uint64_t, int, const void *, void *,
                      int, double, uint64_t, uint64_t, uint64_t, cudaStream_t)
{
  return 0;
}

int fast_try_convolve(Plan *, void *const *, uint32_t, uint32_t, int, double,
                      uint64_t, uint64_t, cudaStream_t)
{
  return 0;
}

}
