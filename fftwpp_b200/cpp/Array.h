/* Array.h -- owning 1-, 2- and 3-index array views with the subset of the
 * reference's Array classes (reference Array.h) that callers of the
 * convolution API use (examples/exampleconv*.cc): construction with an
 * alignment, Nx()/Ny()/Nz()/Size(), a[i][j][k] indexing, implicit conversion
 * to T*, whitespace-separated stream output.  No bounds checking.
 */
#ifndef __Array_h__
#define __Array_h__ 1

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>

namespace Array {

template<class T>
class array1 {
protected:
  T *v;
  size_t size;
  bool owner;
  void take(size_t n, size_t align) {
    void *p=NULL;
    if(align < sizeof(void *)) align=sizeof(void *);
    if(posix_memalign(&p,align,n*sizeof(T) > 0 ? n*sizeof(T) : align)) {
      std::cerr << "\nERROR: Memory limits exceeded." << std::endl;
      exit(1);
    }
    v=(T *) p;
    size=n;
    owner=true;
  }
  array1(const array1&);
  array1& operator=(const array1&);
public:
  array1() : v(NULL), size(0), owner(false) {}
  array1(size_t nx, size_t align=0) {take(nx,align);}
  array1(size_t nx, T *data) : v(data), size(nx), owner(false) {}
  virtual ~array1() {if(owner) free(v);}
  void Dimension(size_t nx, T *data) {
    if(owner) free(v);
    v=data; size=nx; owner=false;
  }
  size_t Nx() const {return size;}
  size_t Size() const {return size;}
  T *operator()() const {return v;}
  operator T*() const {return v;}
  T& operator[](size_t i) const {return v[i];}
  T& operator()(size_t i) const {return v[i];}
  void Load(T a) const {for(size_t i=0; i < size; ++i) v[i]=a;}
  void Load(const T *a) const {memcpy(v,a,size*sizeof(T));}
  array1<T>& operator=(T a) {Load(a); return *this;}
};

template<class T>
class array2 : public array1<T> {
protected:
  size_t nx,ny;
public:
  array2() : nx(0), ny(0) {}
  array2(size_t nx, size_t ny, size_t align=0) : nx(nx), ny(ny) {
    this->take(nx*ny,align);
  }
  array2(size_t nx, size_t ny, T *data) : array1<T>(nx*ny,data), nx(nx), ny(ny) {}
  size_t Nx() const {return nx;}
  size_t Ny() const {return ny;}
  T *operator[](size_t i) const {return this->v+i*ny;}
  T& operator()(size_t i, size_t j) const {return this->v[i*ny+j];}
  T& operator()(size_t i) const {return this->v[i];}
  T *operator()() const {return this->v;}
  array2<T>& operator=(T a) {this->Load(a); return *this;}
};

// row view returned by array3::operator[]
template<class T>
class rows2 {
  T *v;
  size_t nz;
public:
  rows2(T *v, size_t nz) : v(v), nz(nz) {}
  T *operator[](size_t j) const {return v+j*nz;}
  operator T*() const {return v;}
};

template<class T>
class array3 : public array1<T> {
protected:
  size_t nx,ny,nz;
public:
  array3() : nx(0), ny(0), nz(0) {}
  array3(size_t nx, size_t ny, size_t nz, size_t align=0) :
    nx(nx), ny(ny), nz(nz) {
    this->take(nx*ny*nz,align);
  }
  array3(size_t nx, size_t ny, size_t nz, T *data) :
    array1<T>(nx*ny*nz,data), nx(nx), ny(ny), nz(nz) {}
  size_t Nx() const {return nx;}
  size_t Ny() const {return ny;}
  size_t Nz() const {return nz;}
  rows2<T> operator[](size_t i) const {return rows2<T>(this->v+i*ny*nz,nz);}
  T& operator()(size_t i, size_t j, size_t k) const {
    return this->v[(i*ny+j)*nz+k];
  }
  T& operator()(size_t i) const {return this->v[i];}
  T *operator()() const {return this->v;}
  array3<T>& operator=(T a) {this->Load(a); return *this;}
};

template<class T>
std::ostream& operator<<(std::ostream& s, const array1<T>& A)
{
  for(size_t i=0; i < A.Nx(); ++i) s << A()[i] << " ";
  return s;
}

template<class T>
std::ostream& operator<<(std::ostream& s, const array2<T>& A)
{
  const T *p=A();
  for(size_t i=0; i < A.Nx(); ++i) {
    for(size_t j=0; j < A.Ny(); ++j) s << *p++ << " ";
    s << "\n";
  }
  return s << std::flush;
}

template<class T>
std::ostream& operator<<(std::ostream& s, const array3<T>& A)
{
  const T *p=A();
  for(size_t i=0; i < A.Nx(); ++i) {
    for(size_t j=0; j < A.Ny(); ++j) {
      for(size_t k=0; k < A.Nz(); ++k) s << *p++ << " ";
      s << "\n";
    }
    s << "\n";
  }
  return s << std::flush;
}

}

#endif
