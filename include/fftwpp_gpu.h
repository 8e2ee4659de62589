/* fftwpp_gpu.h -- thin C ABI between the host C++ classes (convolve.h mirror)
 * and the hand-written sm_100a CUDA kernels.  Plain pointers and sizes only;
 * no CUDA, torch or C++ types in any signature, so the host layer (and any
 * FFI: ctypes, Fortran bind(C), cgo) compiles without nvcc.
 *
 * What each entry point replaces in the reference (/root/reference):
 *   fftwpp_gpu_plan_create     fftPad*::init(): FFTW plans + Zetaqm/ZetaqmS/
 *                              Zetaqp tables   convolve.cc:511-719,1967-2004,
 *                                              4309-4421,5449-5637,126-140
 *   fftwpp_gpu_forward         (fft->*Forward)(f,F,r,W): pre-twiddle/pad loop
 *                              + fftm->fft     convolve.cc:771-1466 (fftPad),
 *                              2006-3550 (Centered), 4439-5200 (Hermitian),
 *                              5726-6600 (Real); fftw++.h:313,827-896
 *   fftwpp_gpu_backward        (fft->*Backward)(F,f,r,W): ifftm->fft +
 *                              post-twiddle accumulate  convolve.cc:1468-1965,
 *                              2038-4307, 4484-5447, 6602-7483
 *   fftwpp_gpu_convolve        Convolution::convolveRaw residue loop with the
 *                              multiplier fused  convolve.cc:7513-7575,
 *                              convolve.h:1114-1149; multipliers
 *                              convolve.cc:26-110
 *   fftwpp_gpu_scale           Convolution{,2,3}::normalize
 *                              convolve.h:1105-1112,1459-1471,1791-1808
 *   fftwpp_gpu_copy3 +         mpitranspose<Complex>::localize0/1 pack/unpack and
 *   fftwpp_gpu_comm_*          exchange, mpi/mpitranspose.h:632-931 (NCCL)
 *
 * All functions return 0 on success or a negative FFTWPP_GPU_E* code; the C++
 * layer turns failures into the reference's "message on cerr + exit" policy
 * (convolve.h:226-232).  There is NO CPU fallback: without a CUDA device the
 * compute entry points return FFTWPP_GPU_ENODEVICE.
 */
#ifndef FFTWPP_GPU_H
#define FFTWPP_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFTWPP_GPU_OK          0
#define FFTWPP_GPU_ENODEVICE  -1
#define FFTWPP_GPU_EINVAL     -2
#define FFTWPP_GPU_ENOMEM     -3
#define FFTWPP_GPU_ECUDA      -4
#define FFTWPP_GPU_EUNSUPPORTED -5
#define FFTWPP_GPU_ENCCL      -6

/* Kinds of padded FFT (reference classes, convolve.h:471,592,701,805). */
#define FFTWPP_KIND_COMPLEX   0  /* fftPad          */
#define FFTWPP_KIND_CENTERED  1  /* fftPadCentered  */
#define FFTWPP_KIND_HERMITIAN 2  /* fftPadHermitian */
#define FFTWPP_KIND_REAL      3  /* fftPadReal      */

/* Multipliers fused into fftwpp_gpu_convolve (convolve.cc:26-110). */
#define FFTWPP_MULT_NONE        0
#define FFTWPP_MULT_BINARY      1
#define FFTWPP_MULT_REALBINARY  2
#define FFTWPP_MULT_CORRELATION 3

/* Sub-block flags. */
#define FFTWPP_SB_CONJ_OUT 1u  /* store conj(FFT): r2c (sign -1) convention of
                                  the r=0 block of fftPadReal */
#define FFTWPP_SB_SELFCONJ 2u  /* fftPadReal, p > 2: the block holds both halves
                                  of a conjugate-symmetric spectrum (r=0, u=0),
                                  so the backward pass adds Re() once instead of
                                  twice */

/* One FFT sub-block of a residue pass.  The forward pass computes
 *   W[s] = sum_{j in [jmin,jmax), j = s (mod mlen)} zeta_N^{k0*j} * g(j)
 *   out[l] = sum_s zeta_mlen^{l*s} W[s],  l in [0,nout)
 * where g is the logical (origin-shifted / Hermitian-extended / real) input,
 * and stores out[l] at word offset off_* + S_out*l + c of the output buffer.
 * off_call is the reference's own layout for a single forward(r) call
 * (block d at b*d, convolve.h:297-326,956-979); off_all is the layout when
 * every residue is produced by one launch (all blocks concatenated).
 * Offsets are in output words (Complex, or double for Hermitian). */
typedef struct {
  uint32_t mlen;
  uint32_t nout;
  uint32_t flags;
  uint32_t reserved;
  uint64_t k0;
  uint64_t off_call;
  uint64_t off_all;
} fftwpp_gpu_subblock;

typedef struct {
  int32_t kind;      /* FFTWPP_KIND_*                                     */
  int32_t reserved;
  uint64_t L;        /* reference L (logical unpadded length)             */
  uint64_t Lin;      /* stored input words per column (inputLength())     */
  uint64_t N;        /* padded length m*q                                 */
  uint64_t m;        /* inner FFT length                                  */
  uint64_t C;        /* number of interleaved columns                     */
  uint64_t S;        /* stride between successive elements (words)        */
  uint64_t nsub;     /* number of sub-blocks                              */
  const fftwpp_gpu_subblock *sub; /* host array [nsub]                    */
} fftwpp_gpu_pad_desc;

typedef struct fftwpp_gpu_plan fftwpp_gpu_plan;

/* ---- device / memory / stream plumbing ---- */
int fftwpp_gpu_device_count(void);
int fftwpp_gpu_set_device(int device);
int fftwpp_gpu_malloc(void **ptr, size_t bytes);
int fftwpp_gpu_free(void *ptr);
int fftwpp_gpu_malloc_host(void **ptr, size_t bytes);   /* pinned */
int fftwpp_gpu_free_host(void *ptr);
int fftwpp_gpu_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int fftwpp_gpu_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int fftwpp_gpu_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int fftwpp_gpu_memset(void *dst, int value, size_t bytes, void *stream);
/* strided 2-D copies (rows of `width` bytes) */
int fftwpp_gpu_memcpy2d(void *dst, size_t dpitch, const void *src, size_t spitch,
                        size_t width, size_t height, int kind /*0 h2d,1 d2h,2 d2d*/,
                        void *stream);
int fftwpp_gpu_stream_sync(void *stream);
/* streams / events for overlapping the exchange with compute */
int fftwpp_gpu_stream_create(void **stream);
int fftwpp_gpu_stream_destroy(void *stream);
int fftwpp_gpu_event_create(void **event);
int fftwpp_gpu_event_destroy(void *event);
int fftwpp_gpu_event_record(void *event, void *stream);
int fftwpp_gpu_stream_wait_event(void *stream, void *event);
int fftwpp_gpu_event_sync(void *event);
int fftwpp_gpu_device_sync(void);
/* 1 if ptr is device (or managed) memory, 0 if host, <0 on error */
int fftwpp_gpu_is_device_ptr(const void *ptr);
const char *fftwpp_gpu_last_error(void);
/* number of kernels launched by this library since load (bench bookkeeping) */
uint64_t fftwpp_gpu_launch_count(void);

/* ---- plans ---- */
int fftwpp_gpu_plan_create(const fftwpp_gpu_pad_desc *desc, fftwpp_gpu_plan **plan);
int fftwpp_gpu_plan_destroy(fftwpp_gpu_plan *plan);

/* Two-stage ("inner", p > 2; reference fftPad::forwardInner, convolve.cc
 * :1227-1466) transforms: `child` is the length-p pass over m interleaved
 * columns; its outputs (row l of sub-block k0, column c) are multiplied by
 * zeta_N^{(n*l+k0)*c} of `parent` (N = parent padded length) on the way out of
 * fftwpp_gpu_forward and by the conjugate on the way into fftwpp_gpu_backward.
 * `parent` must outlive `child`.  Power-of-two strided passes only. */
int fftwpp_gpu_plan_set_outer(fftwpp_gpu_plan *child, fftwpp_gpu_plan *parent,
                              uint64_t n);

/* ---- optional per-launch timing with CUDA events on the launch stream ----
 * Plans carry a small tag (0..15) naming the pass they serve (the host classes
 * use 1 = x, 2 = y, 3 = z).  While profiling is enabled every launch is
 * bracketed by an event pair; profile_read synchronises the device and returns,
 * for key = 4*tag + op (op 0 forward, 1 backward, 2 fused convolution, 3
 * other), the summed milliseconds and the number of launches (arrays of 64). */
int fftwpp_gpu_plan_set_tag(fftwpp_gpu_plan *plan, int tag);
int fftwpp_gpu_profile_enable(int on);
int fftwpp_gpu_profile_read(double *ms, uint64_t *count);

/* Forward residue pass over sub-blocks [sb0, sb0+nsb) for `nrows` independent
 * rows.  f: input words (Complex, or double for FFTWPP_KIND_REAL); F: output
 * words.  Row r reads f + r*f_rowstride and writes F + r*F_rowstride (strides
 * in words of the respective arrays).  all_layout selects off_all/off_call. */
int fftwpp_gpu_forward(fftwpp_gpu_plan *plan, uint64_t sb0, uint64_t nsb,
                       int all_layout, const void *f, void *F,
                       uint64_t nrows, uint64_t f_rowstride,
                       uint64_t F_rowstride, void *stream);

/* Backward (adjoint) pass: f (=|+=) scale * sum over the given sub-blocks.
 * accumulate != 0 adds to the existing contents of f. */
int fftwpp_gpu_backward(fftwpp_gpu_plan *plan, uint64_t sb0, uint64_t nsb,
                        int all_layout, const void *F, void *f,
                        int accumulate, double scale,
                        uint64_t nrows, uint64_t F_rowstride,
                        uint64_t f_rowstride, void *stream);

/* Fused exchange (replaces mpitranspose localize1/localize0 + the pack/unpack
 * passes): the same passes as fftwpp_gpu_forward/backward in the all-residues
 * layout, but output row r of plane i is stored at
 *   (word *) rowbase[r] + i*rowstride[r] + column
 * where rowbase entries may point into PEER GPUs' memory (mapped with
 * fftwpp_gpu_ipc_open): the transposed data is written over NVLink straight
 * from the FFT epilogue, tile by tile, overlapping transfer and math.
 * forward_mapped: r = all-layout output row; backward_mapped: r = input row j
 * of the transformed dimension, i = plane0 + row index of the batch.
 * rowbase/rowstride are DEVICE arrays. */
int fftwpp_gpu_forward_mapped(fftwpp_gpu_plan *plan, uint64_t sb0,
                              uint64_t nsb, const void *f,
                              const uint64_t *rowbase,
                              const int64_t *rowstride, uint64_t nrows,
                              uint64_t f_rowstride, void *stream);
int fftwpp_gpu_backward_mapped(fftwpp_gpu_plan *plan, uint64_t sb0,
                               uint64_t nsb, const void *F,
                               const uint64_t *rowbase,
                               const int64_t *rowstride, uint64_t plane0,
                               double scale, uint64_t nrows,
                               uint64_t F_rowstride, void *stream);
/* Fused exchange through the TMA unit (preferred over the row maps above when
 * the plan's TMA-staged kernels apply; returns FFTWPP_GPU_EUNSUPPORTED
 * otherwise, and the caller falls back to the *_mapped entry points).  The
 * output rows of the pass are split among `ndest` destinations, destination p
 * owning rows [row0,row0+rows): output row r of launch row (plane) i, column c
 * is stored at
 *   (word *) base + (r-row0)*row_stride + (plane0+i)*plane_stride + c
 * where base may point into a PEER GPU's memory (fftwpp_gpu_ipc_open).  Each
 * destination becomes a tensor map; the kernel's bulk tensor stores
 * (cp.async.bulk.tensor) are clipped to the owner's rows by the tensor
 * bounds, so the transposed data crosses NVLink as TMA traffic, not as
 * per-thread stores.  forward: r = all-residues output row; backward: r =
 * input index j of the transformed dimension.  Strides in 16-byte words.
 * Replaces mpitranspose localize1/localize0 (mpi/mpitranspose.h:632-931). */
typedef struct {
  void *base;
  uint64_t row0;
  uint64_t rows;
  uint64_t row_stride;
  uint64_t plane_stride;
} fftwpp_gpu_dest;
int fftwpp_gpu_forward_dests(fftwpp_gpu_plan *plan, uint64_t sb0, uint64_t nsb,
                             const void *f, const fftwpp_gpu_dest *dests,
                             int ndest, uint64_t plane0, uint64_t nrows,
                             uint64_t f_rowstride, void *stream);
int fftwpp_gpu_backward_dests(fftwpp_gpu_plan *plan, uint64_t sb0,
                              uint64_t nsb, const void *F,
                              const fftwpp_gpu_dest *dests, int ndest,
                              uint64_t plane0, double scale, uint64_t nrows,
                              uint64_t F_rowstride, void *stream);
/* 1 if the plan's forward (backward != 0: backward) pass can run with a row
 * map, 0 if not (no kernel is launched).  Ranks of a distributed convolution
 * agree on the fused exchange by reducing this predicate. */
int fftwpp_gpu_mapped_supported(fftwpp_gpu_plan *plan, int backward);
/* CUDA IPC: export a cudaMalloc'ed buffer / map a peer's buffer (64-byte
 * handles, exchanged by any means) */
int fftwpp_gpu_ipc_get_handle(void *devptr, char *handle64);
int fftwpp_gpu_ipc_open(const char *handle64, void **peerptr);
int fftwpp_gpu_ipc_close(void *peerptr);

/* Fused 1-D convolution of nrows independent rows (plan must have C==1):
 * for every sub-block, forward all A inputs, apply the multiplier, backward
 * the B outputs and accumulate; padded data never leaves the SM.  The result
 * times `scale` overwrites f[0..B).  f: array of max(A,B) device pointers
 * (host array of pointers). */
int fftwpp_gpu_convolve(fftwpp_gpu_plan *plan, void *const *f, uint32_t A,
                        uint32_t B, int mult, double scale, uint64_t nrows,
                        uint64_t rowstride, void *stream);

/* x[i] *= scale over a (n0 x n1 x n2) box of doubles with strides s0,s1 (in
 * doubles) and unit stride in the last dimension. */
int fftwpp_gpu_scale(double *x, double scale, uint64_t n0, uint64_t n1,
                     uint64_t n2, uint64_t s0, uint64_t s1, void *stream);

/* Pack/unpack for the distributed transpose (mpi/mpitranspose.h:632-931):
 * copies an (n0 x n1 x n2)-word box between two strided layouts
 * (16-byte words). dst[i*d0+j*d1+k] = src[i*s0+j*s1+k]. */
int fftwpp_gpu_copy3(void *dst, const void *src, uint64_t n0, uint64_t n1,
                     uint64_t n2, uint64_t d0, uint64_t d1, uint64_t s0,
                     uint64_t s1, void *stream);

/* ---- NCCL exchange for the distributed transpose
 * (replaces mpi/mpitranspose.h:132-161,632-931) ---- */
/* 128-byte NCCL unique id: create on one rank, broadcast by any means */
int fftwpp_gpu_comm_unique_id(char *id128);
int fftwpp_gpu_comm_create(int rank, int size, const char *id128, void **comm);
int fftwpp_gpu_comm_destroy(void *comm);
/* stream-ordered barrier over all ranks, and byte all-gather */
int fftwpp_gpu_comm_barrier(void *comm, void *stream);
int fftwpp_gpu_comm_allgather(void *comm, const void *send, void *recv,
                              uint64_t bytes, void *stream);
int fftwpp_gpu_comm_rank(void *comm);
int fftwpp_gpu_comm_size(void *comm);
/* MPI_Alltoallv semantics, counts and displacements in BYTES */
int fftwpp_gpu_comm_alltoallv(void *comm, const void *send,
                              const uint64_t *scount, const uint64_t *sdispl,
                              void *recv, const uint64_t *rcount,
                              const uint64_t *rdispl, void *stream);

#ifdef __cplusplus
}
#endif

#endif
