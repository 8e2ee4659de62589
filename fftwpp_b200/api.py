"""Python mirror of the reference's wrapper classes (reference wrappers/fftwpp.py)
over the generic handle API of lib_fftwpp.so.

Arrays are numpy arrays (host; staged through the device by the library) or
torch CUDA tensors (device; convolved in place, asynchronously on the current
library stream).
"""
import ctypes

import numpy as np

from ._lib import lib

KIND_COMPLEX, KIND_CENTERED, KIND_HERMITIAN, KIND_REAL = 0, 1, 2, 3
MULT_NONE, MULT_BINARY, MULT_REALBINARY, MULT_CORRELATION = 0, 1, 2, 3
FAMILY_COMPLEX, FAMILY_HERMITIAN, FAMILY_REAL = 0, 1, 2

_INFO = ("L M C S m p q n R dr D D0 l b inplace overwrite centered inputLength "
         "wordSize doubles outputSize workSizeW workSizeV nloops loop2 conjugates "
         "residueBlocks paddedSize normalization repad allRows").split()


def _ptr(a):
    """Raw data pointer of a numpy array or torch tensor."""
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("arrays must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensors must be contiguous")
        return a.data_ptr()
    # any other device array (cupy, numba, ...): the CUDA array interface
    cai = getattr(a, "__cuda_array_interface__", None)
    if cai is not None:
        if cai.get("strides") is not None:
            # strides are given only for non-C-contiguous arrays (interface v2+)
            item = np.dtype(cai["typestr"]).itemsize
            expect, acc = [], item
            for n in reversed(cai["shape"]):
                expect.insert(0, acc)
                acc *= n
            if tuple(cai["strides"]) != tuple(expect):
                raise ValueError("device arrays must be C-contiguous")
        return int(cai["data"][0])
    raise TypeError("expected a numpy array, a torch tensor or an object with "
                    "__cuda_array_interface__")


def launch_count():
    """Number of CUDA kernels this library has launched in this process."""
    return int(lib.fftwpp_gpu_launch_count())


def set_stream(cuda_stream):
    """Use the given cudaStream_t (int) for all subsequent launches."""
    lib.fftwpp_set_stream(ctypes.c_void_p(cuda_stream))


class Pad:
    """One padded FFT (reference fftPad / fftPadCentered / fftPadHermitian /
    fftPadReal, convolve.h:471-980) with the accessors tests/hybrid*.cc walk."""

    def __init__(self, kind, L, M, C=1, S=0, m=0, D=0, I=-1, A=1, B=1, mult=MULT_NONE):
        self.kind = kind
        self._h = lib.fftwpp_pad_create(kind, L, M, C, S, m, D, I, A, B, mult)
        buf = (ctypes.c_size_t * 32)()
        lib.fftwpp_pad_info(self._h, buf)
        self.info = dict(zip(_INFO, [int(v) for v in buf]))
        for k, v in self.info.items():
            setattr(self, k, v)

    def close(self):
        if self._h:
            lib.fftwpp_pad_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def increment(self, r):
        return int(lib.fftwpp_pad_increment(self._h, r))

    def blocksize(self, r):
        return int(lib.fftwpp_pad_blocksize(self._h, r))

    def noutputs(self, r):
        return int(lib.fftwpp_pad_noutputs(self._h, r))

    def span(self, r):
        return int(lib.fftwpp_pad_span(self._h, r))

    def index(self, r, i):
        return int(lib.fftwpp_pad_index(self._h, r, i))

    def residue_calls(self):
        r = 0
        out = []
        while r < self.R:
            out.append(r)
            r += self.increment(r)
        return out

    def forward(self, f, r=0, F=None):
        """fft->forward(f,F,r): returns F (outputSize Complex words)."""
        if F is None:
            F = np.zeros(self.outputSize, dtype=np.complex128)
        lib.fftwpp_pad_forward(self._h, _ptr(f), _ptr(F), r)
        return F

    def backward(self, F, f, r=0):
        """fft->backward(F,f,r): assigns for r == 0, accumulates for r > 0."""
        lib.fftwpp_pad_backward(self._h, _ptr(F), _ptr(f), r)
        return f


class HybridConv:
    """Hybrid dealiased convolution in 1, 2 or 3 dimensions.

    family 0: complex (reference HybridConvolution{,2,3}),
    family 1: centered Hermitian (HybridConvolutionHermitian{,2,3}),
    family 2: real input (tests/hybridconvr{,2,3}.cc object graph).
    """

    def __init__(self, L, M=None, family=FAMILY_COMPLEX, m=None, D=None, I=None,
                 Sx=0, Sy=0, A=2, B=1, mult=None, device_mult=None, indexed=False):
        """mult: one of the MULT_* built-ins (fused on the GPU), or a Python
        callable mult(F, n, r, offset) -- the reference's user `multiplier`
        (convolve.h:78-82): F is a list of max(A,B) numpy views of the n
        transformed words of one residue block, results go to F[0:B].
        device_mult (with a callable mult): device_mult(ptrs, n, r, offset,
        stream) receives the DEVICE addresses of the same blocks and must only
        enqueue GPU work; the transformed data then stays on the GPU.
        indexed=True: the callables receive (F, n, ctx[, stream]) where ctx
        carries the whole transformed multi-index (ctx.outer, ctx.index(j))."""
        L = [int(v) for v in (L if hasattr(L, "__len__") else [L])]
        dim = len(L)
        if M is None:
            if family == FAMILY_HERMITIAN:
                M = [3 * ((l + 1) // 2) - 2 * (l % 2) for l in L]
            else:
                M = [A * l - A + 1 for l in L]
        M = [int(v) for v in (M if hasattr(M, "__len__") else [M])]
        if mult is None:
            mult = MULT_REALBINARY if family == FAMILY_HERMITIAN else MULT_BINARY
        arr = ctypes.c_size_t * dim
        larr = ctypes.c_long * dim
        m = arr(*([0] * dim if m is None else [int(v) for v in m]))
        D = arr(*([0] * dim if D is None else [int(v) for v in D]))
        I = larr(*([-1] * dim if I is None else [int(v) for v in I]))
        self.dim, self.family, self.L, self.M, self.A, self.B = dim, family, L, M, A, B
        if callable(mult):
            self._cb = self._callbacks(mult, device_mult, family, max(A, B), indexed)
            self._h = lib.fftwpp_conv_create_custom(dim, family, arr(*L), arr(*M), m, D, I,
                                                    Sx, Sy, A, B, *self._cb)
        else:
            self._h = lib.fftwpp_conv_create(dim, family, arr(*L), arr(*M), m, D, I,
                                             Sx, Sy, A, B, mult)
        self.doubles = int(lib.fftwpp_conv_doubles(self._h))

    @staticmethod
    def _callbacks(mult, device_mult, family, narrays, indexed=False):
        """ctypes thunks for a Python multiplier pair (kept alive by the object)."""
        import numpy as np
        from ._lib import HOST_MULT, DEVICE_MULT
        dtype = np.float64 if family == FAMILY_HERMITIAN else np.complex128

        class Context(tuple):
            """(r, offset) as before, plus the transformed multi-index of the
            reference's `Indices` (convolve.h:48-76): .outer = indices->index[]
            (outermost dimension last), .index(j) = fft->index(r, j+offset)."""
            outer = ()
            _ind = None

            def index(self, j):
                return int(lib.fftwpp_indices_index(self._ind, j))

        def context(indices):
            r, off = ctypes.c_size_t(), ctypes.c_size_t()
            lib.fftwpp_indices_get(indices, ctypes.byref(r), ctypes.byref(off))
            c = Context((int(r.value), int(off.value)))
            n = int(lib.fftwpp_indices_size(indices))
            c.outer = tuple(int(lib.fftwpp_indices_outer(indices, d)) for d in range(n))
            c._ind = indices
            return c

        def host(F, n, indices, threads):
            views = [np.ctypeslib.as_array(ctypes.cast(F[a], ctypes.POINTER(ctypes.c_double)),
                                           shape=(n * (dtype().itemsize // 8),)).view(dtype)
                     for a in range(narrays)]
            ctx = context(indices)
            if indexed:
                mult(views, n, ctx)
            else:
                mult(views, n, *ctx)

        def device(F, n, indices, stream):
            ctx = context(indices)
            if indexed:
                device_mult([int(F[a]) for a in range(narrays)], n, ctx, stream)
            else:
                device_mult([int(F[a]) for a in range(narrays)], n, *ctx, stream)

        return (HOST_MULT(host),
                DEVICE_MULT(device) if device_mult else ctypes.cast(None, DEVICE_MULT))

    def close(self):
        if self._h:
            lib.fftwpp_conv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def params(self, d):
        buf = (ctypes.c_size_t * 8)()
        lib.fftwpp_conv_params(self._h, d, buf)
        return dict(zip("m p q n D inplace C S".split(), [int(v) for v in buf]))

    def set_plane_chunk(self, chunk):
        lib.fftwpp_conv_set_plane_chunk(self._h, int(chunk))

    def convolve(self, arrays, normalized=True):
        """In-place convolution: the B outputs overwrite arrays[0:B]."""
        n = max(self.A, self.B)
        if len(arrays) < n:
            raise ValueError("need %d arrays" % n)
        ptrs = (ctypes.c_void_p * n)(*[_ptr(a) for a in arrays[:n]])
        lib.fftwpp_conv_convolve(self._h, ptrs, 1 if normalized else 0)
        return arrays[0] if self.B == 1 else arrays[:self.B]

    def convolve_async(self, arrays, slot=0, normalized=True):
        """Pipelined form of convolve() for PINNED host arrays (pinned_array):
        returns at once; wait(slot) blocks until the outputs are back in
        arrays[0:B].  Two slots: alternate them to overlap one convolution's
        PCIe transfers with the other's compute."""
        n = max(self.A, self.B)
        ptrs = (ctypes.c_void_p * n)(*[_ptr(a) for a in arrays[:n]])
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[slot] = arrays  # keep the buffers alive until wait()
        lib.fftwpp_conv_convolve_async(self._h, ptrs, 1 if normalized else 0, slot)

    def wait(self, slot=0):
        lib.fftwpp_conv_wait(self._h, slot)
        if hasattr(self, "_inflight"):
            self._inflight.pop(slot, None)

    def convolve_rows(self, arrays, nrows, rowstride, normalized=True):
        """1-D objects: nrows independent convolutions in one batched launch
        (device tensors shaped (nrows, rowstride))."""
        n = max(self.A, self.B)
        ptrs = (ctypes.c_void_p * n)(*[_ptr(a) for a in arrays[:n]])
        lib.fftwpp_conv_convolve_rows(self._h, ptrs, nrows, rowstride,
                                      1 if normalized else 0)
        return arrays[0]


PROFILE_OPS = ("forward", "backward", "convolve", "other")
PROFILE_PASSES = {0: "-", 1: "x", 2: "y", 3: "z"}


def profile_enable(on=True):
    """Bracket every launch of the library with CUDA events on its stream."""
    lib.fftwpp_gpu_profile_enable(1 if on else 0)


def profile_read():
    """{(pass, op): (total_ms, launches)} since profile_enable(True)."""
    ms = (ctypes.c_double * 64)()
    cnt = (ctypes.c_uint64 * 64)()
    rc = lib.fftwpp_gpu_profile_read(ms, cnt)
    if rc:
        raise RuntimeError(lib.fftwpp_gpu_last_error().decode())
    out = {}
    for k in range(64):
        if cnt[k]:
            out[(PROFILE_PASSES.get(k // 4, str(k // 4)), PROFILE_OPS[k % 4])] = (
                float(ms[k]), int(cnt[k]))
    return out


def pinned_array(shape, dtype):
    """numpy array backed by page-locked host memory (for the e2e timing)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    rc = lib.fftwpp_gpu_malloc_host(ctypes.byref(p), n)
    if rc:
        raise RuntimeError(lib.fftwpp_gpu_last_error().decode())
    buf = (ctypes.c_char * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr
