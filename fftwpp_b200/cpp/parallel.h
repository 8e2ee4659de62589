/* parallel.h -- host-thread helpers with the reference's names (reference
 * parallel.h:30-67).  On the GPU path host threads do no arithmetic; the
 * macros keep caller code that uses them compiling and run the body
 * serially unless the caller itself is built with OpenMP.
 */
#pragma once

#include <cstddef>

#ifdef _OPENMP
#include <omp.h>
#endif

extern size_t threshold;

namespace parallel {

extern size_t lastThreads;

inline size_t get_thread_num()
{
#ifdef _OPENMP
  return omp_get_thread_num();
#else
  return 0;
#endif
}

inline size_t get_thread_num(size_t threads)
{
  return threads > 1 ? get_thread_num() : 0;
}

inline size_t get_max_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void Threshold(size_t threads);

}

#ifdef _OPENMP
#define OMPIF(condition,directive,code)                                   \
  if(threads > 1 && condition) {_Pragma(directive) code} else {code}
#define PARALLEL(code)                                                    \
  if(threads > 1) {_Pragma("omp parallel for num_threads(threads)") code} \
  else {code}
#else
#define OMPIF(condition,directive,code) {code}
#define PARALLEL(code) {code}
#endif

#define PARALLELIF(condition,code)                                        \
  OMPIF(condition,"omp parallel for num_threads(threads)",code)
