/* align.h -- aligned host allocators with the reference's names
 * (reference align.h:109-171: utils::ComplexAlign, doubleAlign, deleteAlign,
 * ceilquotient, align; Array::newAlign/deleteAlign).  The device kernels
 * accept any 16-byte aligned host or device pointer; ALIGNMENT (bytes,
 * reference parallel.cc:18) only shapes the block sizes reported by
 * fftBase, exactly as in the reference.
 */
#ifndef __align_h__
#define __align_h__ 1

#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <new>

#include "Complex.h"

namespace Array {

inline void ArrayExit0(const char *msg)
{
  std::cerr << "\nERROR: " << msg << "." << std::endl;
  exit(1);
}

template<class T>
inline void newAlign(T *&v, size_t len, size_t align)
{
  void *p=NULL;
  if(align < sizeof(void *)) align=sizeof(void *);
  if(align & (align-1)) ArrayExit0("Invalid alignment requested");
  if(posix_memalign(&p,align,len*sizeof(T) > 0 ? len*sizeof(T) : align))
    ArrayExit0("Memory limits exceeded");
  v=(T *) p;
  for(size_t i=0; i < len; ++i) new(v+i) T;
}

template<class T>
inline void deleteAlign(T *v, size_t len)
{
  for(size_t i=len; i > 0;) v[--i].~T();
  free(v);
}

}

namespace utils {

extern size_t ALIGNMENT;

inline size_t ceilquotient(size_t a, size_t b) {return (a+b-1)/b;}

// Round n Complex words up to a multiple of ALIGNMENT bytes.
inline size_t align(size_t n)
{
  return ceilquotient(n*sizeof(Complex),ALIGNMENT)*ALIGNMENT/sizeof(Complex);
}

inline Complex *ComplexAlign(size_t size)
{
  if(size == 0) return NULL;
  Complex *v;
  Array::newAlign(v,size,ALIGNMENT);
  return v;
}

inline double *doubleAlign(size_t size)
{
  double *v;
  Array::newAlign(v,size,ALIGNMENT);
  return v;
}

// n buffers of `size` words carved out of ONE allocation, buffer starts a
// multiple of ALIGNMENT words apart; free with deleteAlign(v[0]); delete [] v;
template<class T>
inline T **alignedBuffers(size_t n, size_t size, T *(*alloc)(size_t))
{
  if(n == 0 || size == 0) return NULL;
  const size_t pitch=ALIGNMENT*ceilquotient(size,ALIGNMENT);
  T *block=alloc((n-1)*pitch+size);
  T **v=new T*[n];
  for(size_t i=0; i < n; ++i) v[i]=block+i*pitch;
  return v;
}

inline Complex **ComplexAlign(size_t n, size_t size)
{
  return alignedBuffers<Complex>(n,size,ComplexAlign);
}

inline double **doubleAlign(size_t n, size_t size)
{
  return alignedBuffers<double>(n,size,doubleAlign);
}

template<class T>
inline void deleteAlign(T *p) {free(p);}

}

#endif
