/* direct_oracle.c -- ORACLE (TEST INFRASTRUCTURE, never linked into or called
 * by the product path; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may use it).
 *
 * Plain-C restatement of the FFT-free direct convolutions that the reference's
 * own tests use as their ground truth (`-E` mode):
 *   linear 1-D            reference tests/direct.h:19-26
 *   centred 1-D           reference tests/direct.h:27-39
 *   linear 2-D / 3-D      reference tests/direct.h:77-88,122-136
 *   Hermitian 1/2/3-D     reference tests/direct.cc:5-98
 * Parity pin: tests/test_oracle.py checks this file against the reference's
 * own direct.h/direct.cc code compiled into oracle/_ref (when built) and
 * against the golden fixtures in tests/golden/.
 *
 * All complex data are interleaved (re,im) doubles, row-major, x slowest.
 */
#include <complex.h>
#include <stddef.h>

typedef double _Complex cplx;

/* h_i = sum_{j<=i} f_j g_{i-j} */
void oracle_direct1_complex(size_t L, const cplx *f, const cplx *g, cplx *h)
{
  for(size_t i=0; i < L; ++i) {
    cplx sum=0.0;
    for(size_t j=0; j <= i; ++j)
      sum += f[j]*g[i-j];
    h[i]=sum;
  }
}

void oracle_direct1_real(size_t L, const double *f, const double *g,
                         double *h)
{
  for(size_t i=0; i < L; ++i) {
    double sum=0.0;
    for(size_t j=0; j <= i; ++j)
      sum += f[j]*g[i-j];
    h[i]=sum;
  }
}

/* Centred data: array index c <-> mode c-floor(L/2); the product of modes
 * a and b lands on mode a+b when that is inside the retained window. */
void oracle_direct1_centered(size_t L, const cplx *f, const cplx *g, cplx *h)
{
  long H=(long) (L/2);
  for(long k=-H; k < (long) L-H; ++k) {
    cplx sum=0.0;
    for(long a=-H; a < (long) L-H; ++a) {
      long b=k-a;
      if(b >= -H && b < (long) L-H)
        sum += f[a+H]*g[b+H];
    }
    h[k+H]=sum;
  }
}

void oracle_direct2_complex(size_t Lx, size_t Ly, const cplx *f,
                            const cplx *g, cplx *h)
{
  for(size_t i=0; i < Lx; ++i)
    for(size_t j=0; j < Ly; ++j) {
      cplx sum=0.0;
      for(size_t k=0; k <= i; ++k)
        for(size_t p=0; p <= j; ++p)
          sum += f[Ly*k+p]*g[Ly*(i-k)+j-p];
      h[Ly*i+j]=sum;
    }
}

void oracle_direct2_real(size_t Lx, size_t Ly, const double *f,
                         const double *g, double *h)
{
  for(size_t i=0; i < Lx; ++i)
    for(size_t j=0; j < Ly; ++j) {
      double sum=0.0;
      for(size_t k=0; k <= i; ++k)
        for(size_t p=0; p <= j; ++p)
          sum += f[Ly*k+p]*g[Ly*(i-k)+j-p];
      h[Ly*i+j]=sum;
    }
}

void oracle_direct3_complex(size_t Lx, size_t Ly, size_t Lz, const cplx *f,
                            const cplx *g, cplx *h)
{
  size_t Sx=Ly*Lz;
  for(size_t i=0; i < Lx; ++i)
    for(size_t j=0; j < Ly; ++j)
      for(size_t k=0; k < Lz; ++k) {
        cplx sum=0.0;
        for(size_t r=0; r <= i; ++r)
          for(size_t p=0; p <= j; ++p)
            for(size_t q=0; q <= k; ++q)
              sum += f[Sx*r+Lz*p+q]*g[Sx*(i-r)+Lz*(j-p)+k-q];
        h[Sx*i+Lz*j+k]=sum;
      }
}

void oracle_direct3_real(size_t Lx, size_t Ly, size_t Lz, const double *f,
                         const double *g, double *h)
{
  size_t Sx=Ly*Lz;
  for(size_t i=0; i < Lx; ++i)
    for(size_t j=0; j < Ly; ++j)
      for(size_t k=0; k < Lz; ++k) {
        double sum=0.0;
        for(size_t r=0; r <= i; ++r)
          for(size_t p=0; p <= j; ++p)
            for(size_t q=0; q <= k; ++q)
              sum += f[Sx*r+Lz*p+q]*g[Sx*(i-r)+Lz*(j-p)+k-q];
        h[Sx*i+Lz*j+k]=sum;
      }
}

/* Hermitian 1-D: f,g hold the H=ceil(L/2) non-negative modes of real
 * signals; modes -j are conj(f_j). */
static cplx herm1(const cplx *f, long j)
{
  return j >= 0 ? f[j] : conj(f[-j]);
}

void oracle_direct1_hermitian(size_t H, const cplx *f, const cplx *g, cplx *h)
{
  long m=(long) H;
  for(long k=0; k < m; ++k) {
    cplx sum=0.0;
    for(long a=1-m; a < m; ++a) {
      long b=k-a;
      if(b > -m && b < m)
        sum += herm1(f,a)*herm1(g,b);
    }
    h[k]=sum;
  }
}

/* Hermitian 2-D: arrays are Lx x Hy (Hy=ceil(Ly/2)), x origin at index
 * x0=floor(Lx/2); x modes run over [-x0, Lx-x0), but for even Lx the most
 * negative (Nyquist) x mode is excluded from the sums, as in the reference
 * (xstart+!xcompact).  Inputs must already be Hermitian-symmetrised. */
void oracle_direct2_hermitian(size_t Lx, size_t Ly, const cplx *f,
                              const cplx *g, cplx *h)
{
  long Hy=(long) ((Ly+1)/2);
  long x0=(long) (Lx/2);
  long xlo=-x0, xhi=(long) Lx-x0;
  long xlo1=(Lx % 2) ? xlo : xlo+1;
  for(long kx=xlo; kx < xhi; ++kx)
    for(long ky=0; ky < Hy; ++ky) {
      cplx sum=0.0;
      for(long px=xlo1; px < xhi; ++px) {
        long qx=kx-px;
        if(qx < xlo1 || qx >= xhi) continue;
        for(long py=1-Hy; py < Hy; ++py) {
          long qy=ky-py;
          if(qy <= -Hy || qy >= Hy) continue;
          cplx fv=py >= 0 ? f[(x0+px)*Hy+py] : conj(f[(x0-px)*Hy-py]);
          cplx gv=qy >= 0 ? g[(x0+qx)*Hy+qy] : conj(g[(x0-qx)*Hy-qy]);
          sum += fv*gv;
        }
      }
      h[(x0+kx)*Hy+ky]=sum;
    }
}

/* Hermitian 3-D: arrays are Lx x Ly x Hz, origins floor(Lx/2), floor(Ly/2). */
void oracle_direct3_hermitian(size_t Lx, size_t Ly, size_t Lz, const cplx *f,
                              const cplx *g, cplx *h)
{
  long Hz=(long) ((Lz+1)/2);
  long x0=(long) (Lx/2), y0=(long) (Ly/2);
  long xlo=-x0, xhi=(long) Lx-x0;
  long ylo=-y0, yhi=(long) Ly-y0;
  long xlo1=(Lx % 2) ? xlo : xlo+1;
  long ylo1=(Ly % 2) ? ylo : ylo+1;
  long Sy=Hz, Sx=(long) Ly*Hz;
  for(long kx=xlo; kx < xhi; ++kx)
    for(long ky=ylo; ky < yhi; ++ky)
      for(long kz=0; kz < Hz; ++kz) {
        cplx sum=0.0;
        for(long px=xlo1; px < xhi; ++px) {
          long qx=kx-px;
          if(qx < xlo1 || qx >= xhi) continue;
          for(long py=ylo1; py < yhi; ++py) {
            long qy=ky-py;
            if(qy < ylo1 || qy >= yhi) continue;
            for(long pz=1-Hz; pz < Hz; ++pz) {
              long qz=kz-pz;
              if(qz <= -Hz || qz >= Hz) continue;
              cplx fv=pz >= 0 ? f[Sx*(x0+px)+Sy*(y0+py)+pz] :
                conj(f[Sx*(x0-px)+Sy*(y0-py)-pz]);
              cplx gv=qz >= 0 ? g[Sx*(x0+qx)+Sy*(y0+qy)+qz] :
                conj(g[Sx*(x0-qx)+Sy*(y0-qy)-qz]);
              sum += fv*gv;
            }
          }
        }
        h[Sx*(x0+kx)+Sy*(y0+ky)+kz]=sum;
      }
}
