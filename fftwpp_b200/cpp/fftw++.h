/* fftw++.h -- what callers of convolve.h see of the reference's FFTW wrapper
 * layer (reference fftw++.h:34-80,230-330): the thread bookkeeping base class
 * and the `fftw` statics (maxthreads, effort) plus the FFTW planner flag
 * names callers OR into fftw::effort.  The FFT classes themselves (fft1d,
 * mfft1d, rcfft1d, ...) are not part of the convolution path: every transform
 * runs inside the sm_100a kernels (csrc/), there are no FFTW plans here.
 */
#ifndef __fftwpp_h__
#define __fftwpp_h__ 1

#include <cstddef>
#include <cstdlib>
#include <iostream>

#include "seconds.h"
#include "parallel.h"
#include "Complex.h"
#include "statistics.h"
#include "align.h"

// FFTW3 public planner flags (api/fftw3.h of FFTW 3.3): accepted, ignored
#ifndef FFTW_MEASURE
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_CONSERVE_MEMORY (1U << 2)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_WISDOM_ONLY (1U << 21)
#define FFTW_NO_SIMD (1U << 17)
#endif

namespace fftwpp {

// Thread bookkeeping kept for source compatibility (reference fftw++.h:59-80)
class ThreadBase {
public:
  size_t threads;
  size_t innerthreads;
  ThreadBase() : threads(1), innerthreads(1) {}
  ThreadBase(size_t threads) : threads(threads), innerthreads(1) {}
  void Threads(size_t nthreads) {threads=nthreads;}
  size_t Threads() {return threads;}
  size_t Innerthreads() {return innerthreads;}
};

// The statics of the reference's fftw base class (reference fftw++.cc:14,17)
class fftw {
public:
  static size_t maxthreads;
  static size_t effort;
};

}

#endif
