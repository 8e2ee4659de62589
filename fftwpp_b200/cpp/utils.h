/* utils.h -- small helpers and the command-line conventions of the
 * reference's test drivers (reference utils.h:8-257): the harness globals
 * that tests/options.cc defines, optionsHybrid/usageHybrid, ceilpow2,
 * ceilpow, the padding helpers.  Declarations only where the definition
 * belongs to the caller's harness (options.cc).
 */
#ifndef __fftwpputils_h__
#define __fftwpputils_h__ 1

#include <cstddef>
#include <cmath>
#include <iostream>
#include <getopt.h>

#include "seconds.h"
#include "Complex.h"

// defined by the test harness (reference tests/options.cc:19-26)
extern double s;  // time limit (seconds)
extern size_t N;  // minimum number of samples
extern size_t C;  // number of padded FFTs computed together
extern size_t S;  // stride between them
extern int stats; // statistic reported by timings()

namespace utils {

template<class T, class U>
inline T max(const T a, const U b) {return a > (T) b ? a : (T) b;}

// x^y by repeated squaring
template<class T>
inline T pow(T x, size_t y)
{
  T r=1;
  for(; y; y >>= 1, x *= x)
    if(y & 1) r *= x;
  return r;
}

// smallest power of two >= n (n > 0)
inline size_t ceilpow2(size_t n)
{
  size_t v=1;
  while(v < n) v <<= 1;
  return v;
}

// smallest power of p >= n
inline size_t ceilpow(size_t p, size_t n)
{
  size_t v=1;
  while(v < n) v *= p;
  return v;
}

inline size_t padding(size_t n)
{
  std::cout << "min padded buffer=" << n << std::endl;
  return ceilpow2(n);
}
inline size_t cpadding(size_t m) {return padding(2*m-1);}
inline size_t hpadding(size_t m) {return padding(3*m-2);}
inline size_t tpadding(size_t m) {return padding(4*m-3);}

// defined in the caller's options.cc (reference tests/options.cc:39-211)
extern void optionsHybrid(int argc, char *argv[], bool fft=false,
                          bool mpi=false);

// option summary of the hybrid test drivers (flags as parsed by
// reference tests/options.cc:80-207)
inline void usageHybrid(bool fft=false, bool mpi=false)
{
  static const char *common[]={
    "-a\t\t accuracy test",
    "-c\t\t use centered tranforms (if possible)",
    "-h\t\t help",
    "-m n\t\t use subtransform size n",
    "-t\t\t show times produced by optimizer",
    NULL};
  std::cerr << "Options: " << std::endl;
  for(const char **p=common; *p; ++p) std::cerr << *p << std::endl;
  if(fft) std::cerr << "-C n\t\t compute n padded FFTs at a time" << std::endl;
  std::cerr << "-D n\t\t number n of blocks to process at a time\n"
            << "-E\t\t compute relative error using direct convolution "
            << "(sets s=0 and forces normalization)\n"
            << "-I\t\t (0=out-of-place, 1=in-place) FFTs "
            << "[by default I=1 only for multiple FFTs]\n"
            << "-O\t\t output result (sets s=0)\n"
            << "-R\t\t show which forward and backward routines are used"
            << std::endl;
  if(mpi)
    std::cerr << "-N n\t\t number of iterations" << std::endl;
  else
    std::cerr << "-N t\t\t minimum number of iterations\n"
              << "-s t\t\t time limit (seconds)" << std::endl;
  std::cerr << "-L n\t\t number n of physical data values\n"
            << "-M n\t\t minimal number n of padded data values" << std::endl;
  if(fft)
    std::cerr << "-S s\t\t use stride s between padded FFTs (defaults to C)"
              << std::endl;
  else
    std::cerr << "-S n\t\t use statistics type n (defaults to 0: MEDIAN)"
              << std::endl;
  std::cerr << "-T n\t\t number n of threads" << std::endl;
}

}

#endif
