/* cfftw++.h -- the include name the reference's C callers use (reference
 * wrappers/cfftw++.h; e.g. wrappers/cexample.c:3).  The declarations live in
 * include/cfftwpp.h (part 1 = the symbols reference wrappers/cfftw++.cc:27-163
 * defines). */
#ifndef CFFTWPP_COMPAT_H
#define CFFTWPP_COMPAT_H
#include "../../include/cfftwpp.h"
#endif
