// fftw3_shim.cc -- independent double-precision FFT behind the FFTW3 API
// subset declared in fftw3.h.  TEST INFRASTRUCTURE ONLY (see fftw3.h): it lets
// the unmodified reference sources under /root/reference be compiled into
// oracle/_ref so the reference's own hybrid-padding logic can be executed as
// an oracle and as the CPU baseline.  Nothing here is derived from FFTW.
//
// Algorithm: Stockham autosort, decimation in frequency, radices 4,2,3,5,7
// and a generic O(r^2) butterfly for any other prime factor; twiddles are
// computed once per plan in long double.  Batched transforms whose batch
// index is the fastest-varying one (idist==1) are processed VB at a time
// through the Stockham stride so strided "many" plans stream whole cache lines.
// Real transforms of even length use the half-length packing; pairs of
// interleaved real transforms are done as one complex transform.

#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef std::complex<double> cplx;

namespace {

int g_plan_threads=1;

struct Engine {
  int n;
  int sign;
  std::vector<int> radices;
  std::vector<cplx> tw; // tw[k]=exp(sign*2*pi*i*k/n)

  Engine(int n, int sign) : n(n), sign(sign), tw(n > 0 ? n : 1) {
    const long double twopi=6.283185307179586476925286766559005768L;
    for(int k=0; k < n; ++k) {
      long double a=twopi*(long double) k/(long double) n;
      tw[k]=cplx((double) cosl(a),(double) (sign*sinl(a)));
    }
    int r=n;
    while(r % 4 == 0) {radices.push_back(4); r /= 4;}
    while(r % 2 == 0) {radices.push_back(2); r /= 2;}
    for(int f=3; f <= r; f += 2)
      while(r % f == 0) {radices.push_back(f); r /= f;}
  }

  // Transform VB interleaved sequences: x[VB*j+v], j < n.  Result is returned
  // in whichever of x,y the last stage wrote; returns that pointer.
  cplx *run(cplx *x, cplx *y, int VB) const {
    int len=n;
    size_t s=VB;
    size_t twstep=1; // n/len
    for(size_t ir=0; ir < radices.size(); ++ir) {
      int r=radices[ir];
      int m=len/r;
      switch(r) {
        case 2: stage2(x,y,m,s,twstep); break;
        case 3: stage3(x,y,m,s,twstep); break;
        case 4: stage4(x,y,m,s,twstep); break;
        case 5: stage5(x,y,m,s,twstep); break;
        default: stageN(x,y,r,m,s,twstep); break;
      }
      cplx *t=x; x=y; y=t;
      len=m;
      s *= r;
      twstep *= r;
    }
    return x;
  }

  inline cplx mulI(cplx a) const { // multiply by sign*i
    return sign > 0 ? cplx(-a.imag(),a.real()) : cplx(a.imag(),-a.real());
  }

  void stage2(const cplx *x, cplx *y, int m, size_t s, size_t twstep) const {
    for(int p=0; p < m; ++p) {
      cplx w=tw[p*twstep];
      const cplx *x0=x+s*p;
      const cplx *x1=x+s*(p+m);
      cplx *y0=y+s*(2*p);
      cplx *y1=y0+s;
      if(p == 0) {
        for(size_t q=0; q < s; ++q) {
          cplx a=x0[q], b=x1[q];
          y0[q]=a+b;
          y1[q]=a-b;
        }
      } else {
        for(size_t q=0; q < s; ++q) {
          cplx a=x0[q], b=x1[q];
          y0[q]=a+b;
          y1[q]=(a-b)*w;
        }
      }
    }
  }

  void stage4(const cplx *x, cplx *y, int m, size_t s, size_t twstep) const {
    for(int p=0; p < m; ++p) {
      cplx w1=tw[p*twstep];
      cplx w2=tw[2*p*twstep];
      cplx w3=tw[3*p*twstep];
      const cplx *x0=x+s*p;
      const cplx *x1=x+s*(p+m);
      const cplx *x2=x+s*(p+2*m);
      const cplx *x3=x+s*(p+3*m);
      cplx *y0=y+s*(4*p);
      cplx *y1=y0+s;
      cplx *y2=y1+s;
      cplx *y3=y2+s;
      for(size_t q=0; q < s; ++q) {
        cplx a=x0[q], b=x1[q], c=x2[q], d=x3[q];
        cplx apc=a+c, amc=a-c, bpd=b+d, jbmd=mulI(b-d);
        y0[q]=apc+bpd;
        y1[q]=(amc+jbmd)*w1;
        y2[q]=(apc-bpd)*w2;
        y3[q]=(amc-jbmd)*w3;
      }
    }
  }

  void stage3(const cplx *x, cplx *y, int m, size_t s, size_t twstep) const {
    const double c=-0.5;
    const double sn=sign*0.86602540378443864676372317075294;
    for(int p=0; p < m; ++p) {
      cplx w1=tw[p*twstep];
      cplx w2=tw[2*p*twstep];
      const cplx *x0=x+s*p;
      const cplx *x1=x+s*(p+m);
      const cplx *x2=x+s*(p+2*m);
      cplx *y0=y+s*(3*p);
      cplx *y1=y0+s;
      cplx *y2=y1+s;
      for(size_t q=0; q < s; ++q) {
        cplx a=x0[q], b=x1[q], d=x2[q];
        cplx t=b+d;
        cplx u=a+c*t;
        cplx v=b-d;
        cplx iv(-sn*v.imag(),sn*v.real());
        y0[q]=a+t;
        y1[q]=(u+iv)*w1;
        y2[q]=(u-iv)*w2;
      }
    }
  }

  void stage5(const cplx *x, cplx *y, int m, size_t s, size_t twstep) const {
    const double c1=0.30901699437494742410229341718282;
    const double c2=-0.80901699437494742410229341718282;
    const double s1=sign*0.95105651629515357211643933337938;
    const double s2=sign*0.58778525229247312916870595463907;
    for(int p=0; p < m; ++p) {
      cplx w1=tw[p*twstep];
      cplx w2=tw[2*p*twstep];
      cplx w3=tw[3*p*twstep];
      cplx w4=tw[4*p*twstep];
      const cplx *x0=x+s*p;
      cplx *y0=y+s*(5*p);
      size_t sm=s*m;
      for(size_t q=0; q < s; ++q) {
        cplx a0=x0[q], a1=x0[q+sm], a2=x0[q+2*sm], a3=x0[q+3*sm],
          a4=x0[q+4*sm];
        cplx t1=a1+a4, t2=a2+a3, t3=a1-a4, t4=a2-a3;
        cplx u1=a0+c1*t1+c2*t2;
        cplx u2=a0+c2*t1+c1*t2;
        cplx v1=s1*t3+s2*t4;
        cplx v2=s2*t3-s1*t4;
        cplx iv1(-v1.imag(),v1.real());
        cplx iv2(-v2.imag(),v2.real());
        y0[q]=a0+t1+t2;
        y0[q+s]=(u1+iv1)*w1;
        y0[q+2*s]=(u2+iv2)*w2;
        y0[q+3*s]=(u2-iv2)*w3;
        y0[q+4*s]=(u1-iv1)*w4;
      }
    }
  }

  void stageN(const cplx *x, cplx *y, int r, int m, size_t s,
              size_t twstep) const {
    size_t rstep=n/r; // tw[rstep*k]=omega_r^k
    std::vector<cplx> a(r);
    for(int p=0; p < m; ++p) {
      for(size_t q=0; q < s; ++q) {
        for(int t=0; t < r; ++t)
          a[t]=x[q+s*(p+(size_t) t*m)];
        for(int u=0; u < r; ++u) {
          cplx sum=a[0];
          for(int t=1; t < r; ++t)
            sum += a[t]*tw[(rstep*(((size_t) t*u) % r)) % n];
          y[q+s*((size_t) r*p+u)]=sum*tw[((size_t) p*u*twstep) % n];
        }
      }
    }
  }
};

enum Kind {C2C, R2C, C2R, COPY};

} // namespace

struct fftw_shim_plan_s {
  Kind kind;
  int n;
  int howmany;
  long istride, idist, ostride, odist;
  int sign;
  int threads;
  Engine *full; // length n, given sign (C2C) or -1/+1 (R2C/C2R)
  Engine *half; // length n/2 for even-length real transforms
  std::vector<cplx> rtw; // exp(-+2 pi i k/n) for the real post/pre-processing
  // COPY (rank-0 guru r2r) description
  int cdims;
  fftw_iodim cd[4];

  fftw_shim_plan_s() : full(NULL), half(NULL), cdims(0) {}
  ~fftw_shim_plan_s() {
    delete full;
    delete half;
  }
};

namespace {

const int VBMAX=8;

fftw_plan make_plan(Kind kind, int n, int howmany, long istride, long idist,
                    long ostride, long odist, int sign)
{
  fftw_plan p=new fftw_shim_plan_s;
  p->kind=kind;
  p->n=n;
  p->howmany=howmany;
  p->istride=istride;
  p->idist=idist;
  p->ostride=ostride;
  p->odist=odist;
  p->sign=sign;
  p->threads=g_plan_threads;
  p->full=new Engine(n,sign);
  if(kind != C2C && n % 2 == 0 && n >= 2) {
    p->half=new Engine(n/2,sign);
    p->rtw.resize(n/2+1);
    const long double twopi=6.283185307179586476925286766559005768L;
    for(int k=0; k <= n/2; ++k) {
      long double a=twopi*(long double) k/(long double) n;
      p->rtw[k]=cplx((double) cosl(a),(double) (sign*sinl(a)));
    }
  }
  return p;
}

struct Scratch {
  std::vector<cplx> a,b;
  void need(size_t n) {
    if(a.size() < n) {a.resize(n); b.resize(n);}
  }
};

Scratch& scratch()
{
  static thread_local Scratch s;
  return s;
}

void exec_c2c(const fftw_plan p, const cplx *in, cplx *out)
{
  const int n=p->n;
  const int H=p->howmany;
  const bool interleaved=(p->idist == 1 && p->odist == 1 && H > 1);
  const int VB=interleaved ? VBMAX : 1;
  const int nblocks=(H+VB-1)/VB;
#pragma omp parallel for num_threads(p->threads) if(p->threads > 1 && nblocks > 1) schedule(static)
  for(int blk=0; blk < nblocks; ++blk) {
    Scratch& S=scratch();
    int h0=blk*VB;
    int vb=std::min(VB,H-h0);
    S.need((size_t) n*vb);
    cplx *a=S.a.data();
    cplx *b=S.b.data();
    const cplx *src=in+h0*p->idist;
    for(int j=0; j < n; ++j) {
      const cplx *sj=src+j*p->istride;
      cplx *aj=a+(size_t) j*vb;
      for(int v=0; v < vb; ++v)
        aj[v]=sj[v*p->idist];
    }
    cplx *res=p->full->run(a,b,vb);
    cplx *dst=out+h0*p->odist;
    for(int j=0; j < n; ++j) {
      cplx *dj=dst+j*p->ostride;
      const cplx *rj=res+(size_t) j*vb;
      for(int v=0; v < vb; ++v)
        dj[v*p->odist]=rj[v];
    }
  }
}

// One real-to-complex transform (sign in p->sign, FFTW uses -1), contiguous
// real input x[0..n) -> X[0..n/2].
void r2c_one(const fftw_plan p, const double *x, cplx *X, Scratch& S)
{
  const int n=p->n;
  if(p->half) {
    const int h=n/2;
    S.need(h+1);
    cplx *a=S.a.data();
    cplx *b=S.b.data();
    for(int j=0; j < h; ++j)
      a[j]=cplx(x[2*j],x[2*j+1]);
    cplx *Z=p->half->run(a,b,1);
    // X[k]=E[k]+w^k O[k]; E[k]=(Z[k]+conj(Z[h-k]))/2, O[k]=(Z[k]-conj(Z[h-k]))/(2i)
    cplx Z0=Z[0];
    X[0]=cplx(Z0.real()+Z0.imag(),0.0);
    X[h]=cplx(Z0.real()-Z0.imag(),0.0);
    for(int k=1; k < h; ++k) {
      cplx zk=Z[k];
      cplx zc=std::conj(Z[h-k]);
      cplx E=0.5*(zk+zc);
      cplx D=0.5*(zk-zc);
      cplx O(D.imag(),-D.real()); // D/i
      X[k]=E+p->rtw[k]*O;
    }
  } else {
    S.need(n);
    cplx *a=S.a.data();
    cplx *b=S.b.data();
    for(int j=0; j < n; ++j)
      a[j]=cplx(x[j],0.0);
    cplx *Z=p->full->run(a,b,1);
    for(int k=0; k <= n/2; ++k)
      X[k]=Z[k];
  }
}

// One complex-to-real transform (sign +1 in FFTW), X[0..n/2] -> x[0..n).
void c2r_one(const fftw_plan p, const cplx *X, double *x, Scratch& S)
{
  const int n=p->n;
  if(p->half) {
    const int h=n/2;
    S.need(h+1);
    cplx *a=S.a.data();
    cplx *b=S.b.data();
    // Z[k]=E[k]+i*O[k] with E[k]=(X[k]+conj(X[h-k])), O[k]=(X[k]-conj(X[h-k]))*w^k
    for(int k=0; k < h; ++k) {
      cplx xk=X[k];
      cplx xc=std::conj(X[h-k]);
      if(k == 0) {xk=cplx(xk.real(),0.0); xc=cplx(xc.real(),0.0);}
      cplx E=xk+xc;
      cplx O=(xk-xc)*p->rtw[k];
      a[k]=E+cplx(-O.imag(),O.real());
    }
    cplx *z=p->half->run(a,b,1);
    for(int j=0; j < h; ++j) {
      x[2*j]=z[j].real();
      x[2*j+1]=z[j].imag();
    }
  } else {
    S.need(n);
    cplx *a=S.a.data();
    cplx *b=S.b.data();
    a[0]=cplx(X[0].real(),0.0);
    for(int k=1; k <= n/2; ++k) {
      a[k]=X[k];
      a[n-k]=std::conj(X[k]);
    }
    cplx *z=p->full->run(a,b,1);
    for(int j=0; j < n; ++j)
      x[j]=z[j].real();
  }
}

void exec_r2c(const fftw_plan p, const double *in, cplx *out)
{
  const int n=p->n;
  const int H=p->howmany;
  const int e=n/2+1;
  const bool interleaved=(p->idist == 1 && p->odist == 1 && H > 1);
  if(interleaved) {
    // Pairs of interleaved real sequences as one complex sequence.
    const int VB=VBMAX; // real sequences per block (even)
    const int nblocks=(H+VB-1)/VB;
#pragma omp parallel for num_threads(p->threads) if(p->threads > 1 && nblocks > 1) schedule(static)
    for(int blk=0; blk < nblocks; ++blk) {
      Scratch& S=scratch();
      int h0=blk*VB;
      int vr=std::min(VB,H-h0); // real sequences in this block
      int vb=(vr+1)/2;          // complex sequences
      S.need((size_t) n*vb);
      cplx *a=S.a.data();
      cplx *b=S.b.data();
      const double *src=in+h0;
      for(int j=0; j < n; ++j) {
        const double *sj=src+j*p->istride;
        cplx *aj=a+(size_t) j*vb;
        for(int v=0; v < vb; ++v) {
          double re=sj[2*v];
          double im=(2*v+1 < vr) ? sj[2*v+1] : 0.0;
          aj[v]=cplx(re,im);
        }
      }
      cplx *Z=p->full->run(a,b,vb);
      cplx *dst=out+h0;
      for(int k=0; k < e; ++k) {
        int kc=(n-k) % n;
        cplx *dk=dst+k*p->ostride;
        const cplx *zk=Z+(size_t) k*vb;
        const cplx *zc=Z+(size_t) kc*vb;
        for(int v=0; v < vb; ++v) {
          cplx A=zk[v];
          cplx Bc=std::conj(zc[v]);
          cplx X0=0.5*(A+Bc);
          cplx D=0.5*(A-Bc);
          dk[2*v]=X0;
          if(2*v+1 < vr)
            dk[2*v+1]=cplx(D.imag(),-D.real());
        }
      }
    }
    return;
  }
#pragma omp parallel for num_threads(p->threads) if(p->threads > 1 && H > 1) schedule(static)
  for(int h=0; h < H; ++h) {
    static thread_local std::vector<double> xr;
    static thread_local std::vector<cplx> Xc;
    Scratch& S=scratch();
    const double *src=in+h*p->idist;
    cplx *dst=out+h*p->odist;
    const double *x=src;
    if(p->istride != 1) {
      if((int) xr.size() < n) xr.resize(n);
      for(int j=0; j < n; ++j) xr[j]=src[j*p->istride];
      x=xr.data();
    }
    if(p->ostride != 1) {
      if((int) Xc.size() < e) Xc.resize(e);
      r2c_one(p,x,Xc.data(),S);
      for(int k=0; k < e; ++k) dst[k*p->ostride]=Xc[k];
    } else {
      if((const void *) x == (const void *) dst) {
        // in-place contiguous: stage through a copy of the input
        if((int) xr.size() < n) xr.resize(n);
        memcpy(xr.data(),x,n*sizeof(double));
        x=xr.data();
      }
      r2c_one(p,x,dst,S);
    }
  }
}

void exec_c2r(const fftw_plan p, const cplx *in, double *out)
{
  const int n=p->n;
  const int H=p->howmany;
  const int e=n/2+1;
  const bool interleaved=(p->idist == 1 && p->odist == 1 && H > 1);
  if(interleaved) {
    const int VB=VBMAX;
    const int nblocks=(H+VB-1)/VB;
#pragma omp parallel for num_threads(p->threads) if(p->threads > 1 && nblocks > 1) schedule(static)
    for(int blk=0; blk < nblocks; ++blk) {
      Scratch& S=scratch();
      int h0=blk*VB;
      int vr=std::min(VB,H-h0);
      int vb=(vr+1)/2;
      S.need((size_t) n*vb);
      cplx *a=S.a.data();
      cplx *b=S.b.data();
      const cplx *src=in+h0;
      // Z=X0+i*X1 on the full spectrum, X[n-k]=conj(X[k]).
      for(int k=0; k < e; ++k) {
        int kc=(n-k) % n;
        const cplx *sk=src+k*p->istride;
        cplx *ak=a+(size_t) k*vb;
        cplx *ac=a+(size_t) kc*vb;
        bool self=(kc == k);
        for(int v=0; v < vb; ++v) {
          cplx X0=sk[2*v];
          cplx X1=(2*v+1 < vr) ? sk[2*v+1] : cplx(0.0,0.0);
          if(self) {X0=cplx(X0.real(),0.0); X1=cplx(X1.real(),0.0);}
          cplx iX1(-X1.imag(),X1.real());
          ak[v]=X0+iX1;
          if(!self) {
            cplx c0=std::conj(X0);
            cplx c1=std::conj(X1);
            ac[v]=c0+cplx(-c1.imag(),c1.real());
          }
        }
      }
      cplx *z=p->full->run(a,b,vb);
      double *dst=out+h0;
      for(int j=0; j < n; ++j) {
        double *dj=dst+j*p->ostride;
        const cplx *zj=z+(size_t) j*vb;
        for(int v=0; v < vb; ++v) {
          dj[2*v]=zj[v].real();
          if(2*v+1 < vr)
            dj[2*v+1]=zj[v].imag();
        }
      }
    }
    return;
  }
#pragma omp parallel for num_threads(p->threads) if(p->threads > 1 && H > 1) schedule(static)
  for(int h=0; h < H; ++h) {
    static thread_local std::vector<double> xr;
    static thread_local std::vector<cplx> Xc;
    Scratch& S=scratch();
    const cplx *src=in+h*p->idist;
    double *dst=out+h*p->odist;
    if((int) Xc.size() < e) Xc.resize(e);
    for(int k=0; k < e; ++k) Xc[k]=src[k*p->istride];
    if(p->ostride != 1) {
      if((int) xr.size() < n) xr.resize(n);
      c2r_one(p,Xc.data(),xr.data(),S);
      for(int j=0; j < n; ++j) dst[j*p->ostride]=xr[j];
    } else
      c2r_one(p,Xc.data(),dst,S);
  }
}

void unsupported(const char *what)
{
  fprintf(stderr,"fftw3 shim: %s is not implemented (not on the convolve.cc path)\n",what);
  exit(1);
}

} // namespace

extern "C" {

int fftw_init_threads(void) {return 1;}
void fftw_plan_with_nthreads(int nthreads) {g_plan_threads=nthreads > 0 ? nthreads : 1;}
void fftw_cleanup_threads(void) {}

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *, fftw_complex *, int sign,
                           unsigned)
{
  return make_plan(C2C,n,1,1,n,1,n,sign);
}

fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *, const int *, int istride,
                             int idist, fftw_complex *, const int *,
                             int ostride, int odist, int sign, unsigned)
{
  if(rank != 1) unsupported("fftw_plan_many_dft with rank != 1");
  return make_plan(C2C,n[0],howmany,istride,idist,ostride,odist,sign);
}

fftw_plan fftw_plan_dft_r2c_1d(int n, double *, fftw_complex *, unsigned)
{
  return make_plan(R2C,n,1,1,n,1,n/2+1,-1);
}

fftw_plan fftw_plan_many_dft_r2c(int rank, const int *n, int howmany,
                                 double *, const int *, int istride, int idist,
                                 fftw_complex *, const int *, int ostride,
                                 int odist, unsigned)
{
  if(rank != 1) unsupported("fftw_plan_many_dft_r2c with rank != 1");
  return make_plan(R2C,n[0],howmany,istride,idist,ostride,odist,-1);
}

fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *, double *, unsigned)
{
  return make_plan(C2R,n,1,1,n/2+1,1,n,1);
}

fftw_plan fftw_plan_many_dft_c2r(int rank, const int *n, int howmany,
                                 fftw_complex *, const int *, int istride,
                                 int idist, double *, const int *, int ostride,
                                 int odist, unsigned)
{
  if(rank != 1) unsupported("fftw_plan_many_dft_c2r with rank != 1");
  return make_plan(C2R,n[0],howmany,istride,idist,ostride,odist,1);
}

fftw_plan fftw_plan_dft_2d(int, int, fftw_complex *, fftw_complex *, int,
                           unsigned)
{unsupported("fftw_plan_dft_2d"); return NULL;}
fftw_plan fftw_plan_dft_3d(int, int, int, fftw_complex *, fftw_complex *, int,
                           unsigned)
{unsupported("fftw_plan_dft_3d"); return NULL;}
fftw_plan fftw_plan_dft_r2c_2d(int, int, double *, fftw_complex *, unsigned)
{unsupported("fftw_plan_dft_r2c_2d"); return NULL;}
fftw_plan fftw_plan_dft_r2c_3d(int, int, int, double *, fftw_complex *,
                               unsigned)
{unsupported("fftw_plan_dft_r2c_3d"); return NULL;}
fftw_plan fftw_plan_dft_c2r_2d(int, int, fftw_complex *, double *, unsigned)
{unsupported("fftw_plan_dft_c2r_2d"); return NULL;}
fftw_plan fftw_plan_dft_c2r_3d(int, int, int, fftw_complex *, double *,
                               unsigned)
{unsupported("fftw_plan_dft_c2r_3d"); return NULL;}

// Only the rank-0 (pure copy/transpose) form used by fftw++.h Transpose.
fftw_plan fftw_plan_guru_r2r(int rank, const fftw_iodim *, int howmany_rank,
                             const fftw_iodim *howmany_dims, double *,
                             double *, const fftw_r2r_kind *, unsigned)
{
  if(rank != 0 || howmany_rank > 4)
    unsupported("fftw_plan_guru_r2r with rank != 0");
  fftw_plan p=new fftw_shim_plan_s;
  p->kind=COPY;
  p->n=0;
  p->howmany=0;
  p->sign=0;
  p->threads=g_plan_threads;
  p->cdims=howmany_rank;
  for(int d=0; d < howmany_rank; ++d)
    p->cd[d]=howmany_dims[d];
  return p;
}

void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out)
{
  exec_c2c(p,(const cplx *) in,(cplx *) out);
}

void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out)
{
  if((void *) in == (void *) out) {
    // in-place: real and complex rows overlay each other, so stage the input
    size_t extent=(size_t) (p->n-1)*p->istride+(size_t) (p->howmany-1)*p->idist+1;
    std::vector<double> tmp(in,in+extent);
    exec_r2c(p,tmp.data(),(cplx *) out);
    return;
  }
  exec_r2c(p,in,(cplx *) out);
}

void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex *in, double *out)
{
  if((void *) in == (void *) out) {
    size_t extent=(size_t) (p->n/2)*p->istride+(size_t) (p->howmany-1)*p->idist+1;
    const cplx *ci=(const cplx *) in;
    std::vector<cplx> tmp(ci,ci+extent);
    exec_c2r(p,tmp.data(),out);
    return;
  }
  exec_c2r(p,(const cplx *) in,out);
}

void fftw_execute_r2r(const fftw_plan p, double *in, double *out)
{
  if(p->kind != COPY) unsupported("fftw_execute_r2r on a non-copy plan");
  size_t total=1;
  for(int d=0; d < p->cdims; ++d) total *= p->cd[d].n;
  std::vector<double> tmp;
  const double *src=in;
  if(in == out) {
    // in-place transpose: stage through a gathered copy
    tmp.resize(total);
    size_t idx[4]={0,0,0,0};
    for(size_t t=0; t < total; ++t) {
      size_t off=0;
      for(int d=0; d < p->cdims; ++d) off += idx[d]*p->cd[d].is;
      tmp[t]=in[off];
      for(int d=p->cdims-1; d >= 0; --d) {
        if(++idx[d] < (size_t) p->cd[d].n) break;
        idx[d]=0;
      }
    }
    size_t jdx[4]={0,0,0,0};
    for(size_t t=0; t < total; ++t) {
      size_t off=0;
      for(int d=0; d < p->cdims; ++d) off += jdx[d]*p->cd[d].os;
      out[off]=tmp[t];
      for(int d=p->cdims-1; d >= 0; --d) {
        if(++jdx[d] < (size_t) p->cd[d].n) break;
        jdx[d]=0;
      }
    }
    return;
  }
  size_t idx[4]={0,0,0,0};
  for(size_t t=0; t < total; ++t) {
    size_t ioff=0, ooff=0;
    for(int d=0; d < p->cdims; ++d) {
      ioff += idx[d]*p->cd[d].is;
      ooff += idx[d]*p->cd[d].os;
    }
    out[ooff]=src[ioff];
    for(int d=p->cdims-1; d >= 0; --d) {
      if(++idx[d] < (size_t) p->cd[d].n) break;
      idx[d]=0;
    }
  }
}

void fftw_destroy_plan(fftw_plan p) {delete p;}

int fftw_import_wisdom_from_string(const char *) {return 1;}

char *fftw_export_wisdom_to_string(void)
{
  char *s=(char *) malloc(1);
  s[0]=0;
  return s;
}

void fftw_free(void *p) {free(p);}
void *fftw_malloc(size_t n) {return malloc(n);}

}
