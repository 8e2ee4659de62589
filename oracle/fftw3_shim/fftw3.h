/* fftw3.h -- FFTW3-API shim used ONLY to compile the unmodified reference
 * (/root/reference) into oracle/_ref.  TEST INFRASTRUCTURE, not product code.
 *
 * FFTW3 itself is not vendored by the reference and is not installed in this
 * image (reference CMakeLists.txt:17-19 find_library(fftw3)).  This header
 * declares the small subset of the FFTW3 API the reference's fftw++.h uses
 * (fftw++.h:313,477,494,570-573,704-736,927-933,967-973; fftw++.cc:41,48,54)
 * and fftw3_shim.cc backs it with an independent double-precision mixed-radix
 * FFT.  Results are therefore "FFTW++ on an own-FFT shim", never "FFTW++ on
 * FFTW3"; only the leaf DFT differs, all hybrid-padding logic is the
 * reference's own code.
 */
#ifndef FFTWPP_B200_FFTW3_SHIM_H
#define FFTWPP_B200_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct fftw_shim_plan_s *fftw_plan;

typedef struct {
  int n;
  int is;
  int os;
} fftw_iodim;

typedef enum {
  FFTW_R2HC=0, FFTW_HC2R=1, FFTW_DHT=2, FFTW_REDFT00=3, FFTW_REDFT01=4,
  FFTW_REDFT10=5, FFTW_REDFT11=6, FFTW_RODFT00=7, FFTW_RODFT01=8,
  FFTW_RODFT10=9, FFTW_RODFT11=10
} fftw_r2r_kind;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)

#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_CONSERVE_MEMORY (1U << 2)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_NO_SIMD (1U << 17)
#define FFTW_WISDOM_ONLY (1U << 21)

int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
void fftw_cleanup_threads(void);

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out,
                           int sign, unsigned flags);
fftw_plan fftw_plan_dft_2d(int n0, int n1, fftw_complex *in,
                           fftw_complex *out, int sign, unsigned flags);
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in,
                           fftw_complex *out, int sign, unsigned flags);
fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed,
                             int istride, int idist,
                             fftw_complex *out, const int *onembed,
                             int ostride, int odist,
                             int sign, unsigned flags);

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out,
                               unsigned flags);
fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double *in, fftw_complex *out,
                               unsigned flags);
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in,
                               fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_many_dft_r2c(int rank, const int *n, int howmany,
                                 double *in, const int *inembed,
                                 int istride, int idist,
                                 fftw_complex *out, const int *onembed,
                                 int ostride, int odist, unsigned flags);

fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out,
                               unsigned flags);
fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex *in, double *out,
                               unsigned flags);
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex *in,
                               double *out, unsigned flags);
fftw_plan fftw_plan_many_dft_c2r(int rank, const int *n, int howmany,
                                 fftw_complex *in, const int *inembed,
                                 int istride, int idist,
                                 double *out, const int *onembed,
                                 int ostride, int odist, unsigned flags);

fftw_plan fftw_plan_guru_r2r(int rank, const fftw_iodim *dims,
                             int howmany_rank, const fftw_iodim *howmany_dims,
                             double *in, double *out,
                             const fftw_r2r_kind *kind, unsigned flags);

void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out);
void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out);
void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex *in, double *out);
void fftw_execute_r2r(const fftw_plan p, double *in, double *out);

void fftw_destroy_plan(fftw_plan p);

int fftw_import_wisdom_from_string(const char *input_string);
char *fftw_export_wisdom_to_string(void);
void fftw_free(void *p);
void *fftw_malloc(size_t n);

#ifdef __cplusplus
}
#endif

#endif
