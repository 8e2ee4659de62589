/* cfftwpp.h -- C-callable API of lib_fftwpp.so (B200 build).
 *
 * Part 1 keeps the symbols that the reference's wrapper library actually
 * defines (reference wrappers/cfftw++.cc:27-163; declarations
 * wrappers/cfftw++.h), with the same argument meaning, so wrappers/fftwpp.py
 * (ctypes), wrappers/fftwpp.f90 (bind(C)) and wrappers/cexample.c keep
 * working against this library.  Arrays may be host pointers (staged) or
 * device pointers (in place).  The *_dot/_dotf/_work prototypes of the
 * reference header have no definitions in the reference either
 *  and are not provided.
 *
 * Part 2 is a generic handle API over the same host classes
 * (the fftPad family and Convolution, Convolution2, Convolution3) used by the parity tests, bench.py and the
 * distributed driver: explicit (m,D,I), strides, real-data convolutions,
 * residue-level forward/backward and the size/index queries that
 * tests/hybrid*.cc walk.
 */
#ifndef CFFTWPP_B200_H
#define CFFTWPP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#else
#include <complex.h>
#endif

/* interleaved (re,im) doubles, layout-compatible with C99 double _Complex,
 * std::complex<double> and the reference's Complex (Complex.h:29-38) */
#ifdef __cplusplus
typedef __complex__ double fftwpp_cplx;
#else
typedef double _Complex fftwpp_cplx;
#endif

/* ---------------- Part 1: reference wrapper API ---------------- */

/* wrappers/cfftw++.cc:27-43 */
double *create_doubleAlign(size_t n);
void delete_doubleAlign(double *p);
fftwpp_cplx *create_complexAlign(size_t n);
void delete_complexAlign(fftwpp_cplx *p);

/* wrappers/cfftw++.cc:45-52 */
size_t get_fftwpp_maxthreads(void);
void set_fftwpp_maxthreads(size_t nthreads);

/* 1d complex, wrappers/cfftw++.cc:54-66 */
typedef struct HybridConvolution HybridConvolution;
HybridConvolution *fftwpp_create_conv1d(size_t L);
void fftwpp_conv1d_delete(HybridConvolution *conv);
void fftwpp_conv1d_convolve(HybridConvolution *conv, fftwpp_cplx *a,
                            fftwpp_cplx *b);

/* 1d Hermitian, wrappers/cfftw++.cc:68-86 */
typedef struct HybridConvolutionHermitian HybridConvolutionHermitian;
HybridConvolutionHermitian *fftwpp_create_hconv1d(size_t L);
void fftwpp_hconv1d_delete(HybridConvolutionHermitian *conv);
void fftwpp_HermitianSymmetrize(fftwpp_cplx *f);
void fftwpp_hconv1d_convolve(HybridConvolutionHermitian *conv, fftwpp_cplx *a,
                             fftwpp_cplx *b);

/* 2d, wrappers/cfftw++.cc:88-124 */
typedef struct HybridConvolution2 HybridConvolution2;
HybridConvolution2 *fftwpp_create_conv2d(size_t Lx, size_t Ly);
void fftwpp_conv2d_delete(HybridConvolution2 *conv);
void fftwpp_HermitianSymmetrizeX(size_t Hx, size_t Hy, size_t x0,
                                 fftwpp_cplx *f);
void fftwpp_conv2d_convolve(HybridConvolution2 *conv, fftwpp_cplx *a,
                            fftwpp_cplx *b);
typedef struct HybridConvolutionHermitian2 HybridConvolutionHermitian2;
HybridConvolutionHermitian2 *fftwpp_create_hconv2d(size_t Lx, size_t Ly);
void fftwpp_hconv2d_delete(HybridConvolutionHermitian2 *conv);
void fftwpp_hconv2d_convolve(HybridConvolutionHermitian2 *conv, fftwpp_cplx *a,
                             fftwpp_cplx *b);

/* 3d, wrappers/cfftw++.cc:126-163 */
typedef struct HybridConvolution3 HybridConvolution3;
HybridConvolution3 *fftwpp_create_conv3d(size_t Lx, size_t Ly, size_t Lz);
void fftwpp_conv3d_delete(HybridConvolution3 *conv);
void fftwpp_HermitianSymmetrizeXY(size_t Hx, size_t Hy, size_t Hz, size_t x0,
                                  size_t y0, fftwpp_cplx *f);
void fftwpp_conv3d_convolve(HybridConvolution3 *conv, fftwpp_cplx *a,
                            fftwpp_cplx *b);
typedef struct HybridConvolutionHermitian3 HybridConvolutionHermitian3;
HybridConvolutionHermitian3 *fftwpp_create_hconv3d(size_t Lx, size_t Ly,
                                                   size_t Lz);
void fftwpp_hconv3d_delete(HybridConvolutionHermitian3 *conv);
void fftwpp_hconv3d_convolve(HybridConvolutionHermitian3 *conv, fftwpp_cplx *a,
                             fftwpp_cplx *b);

/* ---------------- Part 2: generic handle API ---------------- */

/* kind: 0 fftPad, 1 fftPadCentered, 2 fftPadHermitian, 3 fftPadReal
 * (convolve.h:471,592,701,805).  m == 0: the chooser picks (m,D,I).
 * mult: 0 multNone, 1 multBinary, 2 realMultBinary, 3 multcorrelation. */
void *fftwpp_pad_create(int kind, size_t L, size_t M, size_t C, size_t S,
                        size_t m, size_t D, long I, size_t A, size_t B,
                        int mult);
void fftwpp_pad_destroy(void *pad);
/* out[0..31]: L,M,C,S,m,p,q,n,R,dr,D,D0,l,b,inplace,overwrite,centered,
 * inputLength,wordSize,doubles,outputSize,workSizeW,workSizeV,nloops,loop2,
 * conjugates,residueBlocks,paddedSize,normalization,repad,allRows,0 */
void fftwpp_pad_info(void *pad, size_t *out);
size_t fftwpp_pad_increment(void *pad, size_t r);
size_t fftwpp_pad_blocksize(void *pad, size_t r);
size_t fftwpp_pad_noutputs(void *pad, size_t r);
size_t fftwpp_pad_span(void *pad, size_t r);
size_t fftwpp_pad_index(void *pad, size_t r, size_t i);
/* fft->forward(f,F,r) / fft->backward(F,f,r), host or device pointers */
void fftwpp_pad_forward(void *pad, const double *f, double *F, size_t r);
void fftwpp_pad_backward(void *pad, const double *F, double *f, size_t r);
/* All-residues pass (device pointers) with the output rows split among nsplit
 * owners of ceil-split row ranges inside the same dense buffer: exercises the
 * destination sets of the fused exchange (fftwpp_gpu_forward_dests) on one
 * GPU.  Returns 0, or FFTWPP_GPU_EUNSUPPORTED (-5) if the plan has no
 * TMA-staged kernel. */
int fftwpp_pad_forward_split(void *pad, const void *f, void *F, size_t nsplit);
int fftwpp_pad_backward_split(void *pad, const void *F, void *f, size_t nsplit,
                              double scale);

/* family: 0 complex, 1 centered Hermitian, 2 real (first dimension real).
 * L,M,m,D,I: arrays of `dim` entries in x,y,z order; m[d]==0: chooser.
 * Built like tests/hybridconv{,h,r}{,2,3}.cc build their objects. */
void *fftwpp_conv_create(int dim, int family, const size_t *L, const size_t *M,
                         const size_t *m, const size_t *D, const long *I,
                         size_t Sx, size_t Sy, size_t A, size_t B, int mult);

/* User multipliers (the reference's `multiplier` callback, convolve.h:78-82,
 * passed to Application): F[a] points at the n transformed words (complex, or
 * doubles for Hermitian families) of one residue block of array a; results go
 * to F[0..B).  `host` runs on the CPU between a GPU forward and a GPU backward
 * pass.  `device` (optional) receives DEVICE pointers instead and must only
 * enqueue work on `stream` (a cudaStream_t): the transformed data then never
 * leaves the GPU.  fftwpp_indices_get reads the residue context. */
typedef void fftwpp_multiplier(double **F, size_t n, void *indices,
                               size_t threads);
typedef void fftwpp_device_multiplier(double **F, size_t n, void *indices,
                                      void *stream);
void *fftwpp_conv_create_custom(int dim, int family, const size_t *L,
                                const size_t *M, const size_t *m,
                                const size_t *D, const long *I, size_t Sx,
                                size_t Sy, size_t A, size_t B,
                                fftwpp_multiplier *host,
                                fftwpp_device_multiplier *device);
void fftwpp_indices_get(void *indices, size_t *r, size_t *offset);
/* The transformed multi-index the reference hands to multipliers through
 * `Indices` (convolve.h:48-76; use shown at convolve.cc:39-48): size = number
 * of OUTER dimensions (0 in 1-D, 1 in 2-D, 2 in 3-D); outer(d) = indices->
 * index[d], set per transformed row by each outer level (convolve.h:1442,1759;
 * d = size-1 is the outermost, x); index(j) = fft->index(r,j+offset), the
 * transformed index of element j of this block in the innermost dimension. */
size_t fftwpp_indices_size(void *indices);
size_t fftwpp_indices_outer(void *indices, size_t d);
size_t fftwpp_indices_index(void *indices, size_t j);
void fftwpp_conv_destroy(void *conv);
/* out = {m,p,q,n,D,inplace,C,S} of dimension d */
void fftwpp_conv_params(void *conv, int d, size_t *out);
/* doubles per input array */
size_t fftwpp_conv_doubles(void *conv);
/* f: array of max(A,B) pointers (all host or all device).  normalized != 0:
 * convolve(); else convolveRaw().  Result overwrites f[0..B). */
void fftwpp_conv_convolve(void *conv, double **f, int normalized);
/* 1-D objects only: convolve `nrows` independent rows in one batched launch
 * (row i of array a at f[a]+i*rowstride words; device pointers).  This is what
 * Convolution2/3 issue for their innermost dimension and what the reference's
 * OpenMP loop over rows does (convolve.h:1434-1445); BASELINE config 5. */
/* Pipelined host-buffer entry: enqueue the H2D copy of the inputs, the
 * convolution and the D2H copy of the B outputs on three streams and return at
 * once; fftwpp_conv_wait(slot) blocks until the outputs of that slot have
 * landed in f[0..B).  Two slots (0, 1): while one convolution runs, the next
 * one's inputs are already crossing PCIe and the previous one's outputs are on
 * their way back.  f must be pinned host memory (fftwpp_gpu_malloc_host) that
 * stays untouched until the wait. */
void fftwpp_conv_convolve_async(void *conv, double **f, int normalized, int slot);
void fftwpp_conv_wait(void *conv, int slot);
void fftwpp_conv_convolve_rows(void *conv, double **f, size_t nrows,
                               size_t rowstride, int normalized);
/* batch size (x rows) of the y/z sweep of a 3-D convolution; 0 = all rows */
void fftwpp_conv_set_plane_chunk(void *conv, size_t chunk);

/* ---- distributed (slab over y) 3-D convolution over NCCL: the counterpart
 * of the reference's Convolution3MPI (mpi/mpiconvolve.h:182-305) built as in
 * mpi/tests/hybridconv{,r}3.cc.  comm: handle from fftwpp_gpu_comm_create.
 * family 0 (complex), 1 (centred Hermitian: last dimension holds the
 * ceil(Lz/2) non-negative modes) or 2 (real).  Arrays are the LOCAL slabs
 * Lx x y x Lz (device pointers). ---- */
void *fftwpp_mpiconv3_create(int family, const size_t *L, const size_t *M,
                             const size_t *m, const size_t *D, const long *I,
                             size_t A, size_t B, int mult, int rank, int size,
                             void *comm);
/* Pencil decomposition (reference mpi/mpiconvolve.h:208-216, process grid
 * mpi/mpigroup.h:33-50): y split over the first communicator, z over the
 * second; arrays are the local pencils Lx x y x z.  Families 0 and 2. */
void *fftwpp_mpiconv3_create_pencil(int family, const size_t *L,
                                    const size_t *M, const size_t *m,
                                    const size_t *D, const long *I, size_t A,
                                    size_t B, int mult, int rankY, int sizeY,
                                    void *commY, int rankZ, int sizeZ,
                                    void *commZ);
void fftwpp_mpiconv3_destroy(void *conv);
/* out = {X,Y,Z,x,y,z,x0,y0,z0} (split3) */
void fftwpp_mpiconv3_split(void *conv, size_t *out);
void fftwpp_mpiconv3_params(void *conv, int d, size_t *out);
void fftwpp_mpiconv3_convolve(void *conv, double **f, int normalized);
/* Pipelined form of fftwpp_mpiconv3_convolve for PINNED HOST slabs (one per
 * rank, the rank's own y slice): returns at once; fftwpp_mpiconv3_wait(conv,
 * slot) blocks until the outputs are back in f[0..B).  Two slots; alternate
 * them to overlap one convolution's PCIe transfers with the other's compute.
 * Collective: every rank calls it in the same order. */
void fftwpp_mpiconv3_convolve_async(void *conv, double **f, int normalized,
                                    int slot);
void fftwpp_mpiconv3_wait(void *conv, int slot);
/* byte counts/displacements of exchange `direction` (0 forward, 1 backward);
 * arrays of `size` entries */
void fftwpp_mpiconv3_exchange_table(void *conv, int direction,
                                    unsigned long long *scount,
                                    unsigned long long *sdispl,
                                    unsigned long long *rcount,
                                    unsigned long long *rdispl);
void fftwpp_mpiconv3_set_plane_chunk(void *conv, size_t chunk);
/* family 1: enforce Hermitian symmetry on the distributed data (device slab
 * Lx x y x ceil(Lz/2)); the reference's HermitianSymmetrizeXY(split3&, f),
 * mpi/mpiconvolve.cc:11-142.  Collective over the communicator. */
void fftwpp_mpiconv3_symmetrize(void *conv, double *f);

/* ---- distributed 2-D convolution: the reference's Convolution2MPI
 * (mpi/mpiconvolve.h:72-179; driver mpi/tests/hybridconv2.cc).  Arrays are the
 * LOCAL slabs Lx x y of complex words (family 0), or Lx x (slice of the
 * ceil(Ly/2) stored modes) for the centred Hermitian family 1 (reference
 * mpi/tests/hybridconvh2.cc); L, M, m, D, I have two entries.  Same conventions as the 3-D handle. ---- */
void *fftwpp_mpiconv2_create(int family, const size_t *L, const size_t *M,
                             const size_t *m, const size_t *D, const long *I,
                             size_t A, size_t B, int mult, int rank, int size,
                             void *comm);
void fftwpp_mpiconv2_destroy(void *conv);
void fftwpp_mpiconv2_split(void *conv, size_t *out);
void fftwpp_mpiconv2_params(void *conv, int d, size_t *out);
void fftwpp_mpiconv2_convolve(void *conv, double **f, int normalized);
void fftwpp_mpiconv2_exchange_table(void *conv, int direction,
                                    unsigned long long *scount,
                                    unsigned long long *sdispl,
                                    unsigned long long *rcount,
                                    unsigned long long *rdispl);

/* ---- distributed FFTs: the reference's fft2dMPI / fft3dMPI / rcfft2dMPI /
 * rcfft3dMPI (mpi/mpifftw++.h:37-585) on the same exchange and kernels.
 * kind: 0 complex, 1 real-to-complex.  dims = 2 or 3; N = global extents
 * (real extents for kind 1).  Complex data: x x Y [x Z] in (this rank's x
 * rows), X x y [x Z] out (this rank's y rows); kind 1 halves the last
 * dimension (N/2+1 complex words) and, in 2-D, splits that halved dimension.
 * Arrays are DEVICE pointers; complex arrays hold fftwpp_mpifft_words() words.
 * split: X,Y,Z,x,y,z,x0,y0,z0 of the COMPLEX data. ---- */
void *fftwpp_mpifft_create(int kind, int dims, const size_t *N, int sign,
                           int rank, int size, void *comm);
void fftwpp_mpifft_destroy(void *fft);
void fftwpp_mpifft_split(void *fft, size_t *out);
size_t fftwpp_mpifft_words(void *fft);
/* byte counts / displacements of the exchange (direction 1: the forward
 * transform's x x Y -> X x y, 0: its inverse), for tests; no GPU needed */
void fftwpp_mpifft_exchange_table(void *fft, int direction,
                                  unsigned long long *scount,
                                  unsigned long long *sdispl,
                                  unsigned long long *rcount,
                                  unsigned long long *rdispl);
/* complex: out may be NULL (in place); real: in = double array, out complex */
void fftwpp_mpifft_forward(void *fft, void *in, void *out);
/* complex: out may be NULL; real: in complex (overwritten), out doubles */
void fftwpp_mpifft_backward(void *fft, void *in, void *out);
/* kind 1 only: multiply the real x x Y [x Z] data by (-1)^x (3-D: (-1)^(x+y)),
 * which centres the Fourier origin (reference Shift / Forward0 / Backward0);
 * zero the Nyquist modes of the transformed data (reference deNyquist) */
void fftwpp_mpifft_shift(void *fft, double *f);
void fftwpp_mpifft_denyquist(void *fft, void *f);
/* divide the x x Y [x Z] data (complex kind 0, real kind 1) by the point count */
void fftwpp_mpifft_normalize(void *fft, void *f);

/* stream used by every launch issued through this API (a cudaStream_t) */
void fftwpp_set_stream(void *stream);

#ifdef __cplusplus
}
#endif

#endif
