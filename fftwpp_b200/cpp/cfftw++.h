/* cfftw++.h -- the include name the reference's C and C++ callers use
 * (reference wrappers/cfftw++.h; e.g. wrappers/cexample.c:3,
 * wrappers/cfftw++.cc:14).  The declarations live in include/cfftwpp.h (part 1
 * = the symbols reference wrappers/cfftw++.cc:27-163 defines).  As in the
 * reference header (wrappers/cfftw++.h:13-15), C++ translation units see them
 * inside namespace fftwpp, where the handle types name the bundle classes of
 * HybridConvolution.h. */
#ifndef CFFTWPP_COMPAT_H
#define CFFTWPP_COMPAT_H
#ifdef __cplusplus
#include <stddef.h>
namespace fftwpp {
#endif
#include "../../include/cfftwpp.h"
#ifdef __cplusplus
}
#endif
#endif
