"""ctypes binding to the reference itself (oracle/_ref/libfftwpp_ref.so).

TEST INFRASTRUCTURE: the unmodified /root/reference sources built by
oracle/Makefile against oracle/fftw3_shim (own FFT leaf; FFTW3 is not
installed).  Used as the checker and as the CPU baseline only.
"""
import ctypes
import os

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
path = os.path.join(_here, "_ref", "libfftwpp_ref.so")

c_size_t, c_void_p, c_int, c_long = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_long
P = ctypes.POINTER

_INFO = ("L M C S m p q n R dr D D0 l b inplace overwrite centered inputLength "
         "wordSize doubles outputSize workSizeW workSizeV nloops loop2 conjugates "
         "residueBlocks paddedSize normalization repad").split()


def available():
    return os.path.exists(path)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError("oracle/_ref is not built (make -C oracle ref)")
        L = ctypes.CDLL(path)

        def sig(name, restype, *args):
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = list(args)

        sig("ref_set_maxthreads", None, c_size_t)
        sig("ref_get_max_threads", c_size_t)
        sig("ref_pad_create", c_void_p, c_int, c_size_t, c_size_t, c_size_t, c_size_t,
            c_size_t, c_size_t, c_long, c_size_t, c_size_t, c_int, c_size_t)
        sig("ref_pad_destroy", None, c_void_p)
        sig("ref_pad_info", None, c_void_p, P(c_size_t))
        for n in ("increment", "blocksize", "noutputs", "span"):
            sig("ref_pad_" + n, c_size_t, c_void_p, c_size_t)
        sig("ref_pad_index", c_size_t, c_void_p, c_size_t, c_size_t)
        sig("ref_pad_forward", None, c_void_p, c_void_p, c_void_p, c_size_t)
        sig("ref_pad_backward", None, c_void_p, c_void_p, c_void_p, c_size_t)
        sig("ref_conv_create", c_void_p, c_int, c_int, P(c_size_t), P(c_size_t),
            P(c_size_t), P(c_size_t), P(c_long), c_size_t, c_size_t, c_size_t,
            c_size_t, c_int, c_size_t, c_int)
        sig("ref_conv_destroy", None, c_void_p)
        sig("ref_conv_params", None, c_void_p, c_int, P(c_size_t))
        sig("ref_conv_doubles", c_size_t, c_void_p)
        sig("ref_conv_convolve", None, c_void_p, P(c_void_p), c_int)
        sig("ref_conv_time", None, c_void_p, c_size_t, P(ctypes.c_double))
        sig("ref_direct_complex", None, c_int, P(c_size_t), c_void_p, c_void_p, c_void_p)
        sig("ref_direct_centered1", None, c_size_t, c_void_p, c_void_p, c_void_p)
        sig("ref_direct_real", None, c_int, P(c_size_t), c_void_p, c_void_p, c_void_p)
        sig("ref_direct_hermitian", None, c_int, P(c_size_t), c_void_p, c_void_p, c_void_p)
        sig("ref_symmetrize", None, c_int, P(c_size_t), c_void_p)
        _lib = L
    return _lib


class RefPad:
    """The reference's fftPad* object with forced (m,D,I)."""

    def __init__(self, kind, L, M, C=1, S=0, m=0, D=0, I=-1, A=1, B=1, mult=0, threads=1):
        self._h = lib().ref_pad_create(kind, L, M, C, S, m, D, I, A, B, mult, threads)
        buf = (c_size_t * 32)()
        lib().ref_pad_info(self._h, buf)
        self.info = dict(zip(_INFO, [int(v) for v in buf]))
        for k, v in self.info.items():
            setattr(self, k, v)

    def close(self):
        if self._h:
            lib().ref_pad_destroy(self._h)
            self._h = None

    def increment(self, r):
        return int(lib().ref_pad_increment(self._h, r))

    def blocksize(self, r):
        return int(lib().ref_pad_blocksize(self._h, r))

    def noutputs(self, r):
        return int(lib().ref_pad_noutputs(self._h, r))

    def span(self, r):
        return int(lib().ref_pad_span(self._h, r))

    def index(self, r, i):
        return int(lib().ref_pad_index(self._h, r, i))

    def residue_calls(self):
        r, out = 0, []
        while r < self.R:
            out.append(r)
            r += self.increment(r)
        return out

    def forward(self, f, r=0):
        F = np.zeros(self.outputSize, dtype=np.complex128)
        lib().ref_pad_forward(self._h, f.ctypes.data, F.ctypes.data, r)
        return F

    def backward(self, F, f, r=0):
        lib().ref_pad_backward(self._h, F.ctypes.data, f.ctypes.data, r)
        return f


class RefConv:
    """The reference's Convolution{,2,3} built like tests/hybridconv*.cc."""

    def __init__(self, L, M, family=0, m=None, D=None, I=None, Sx=0, Sy=0, A=2, B=1,
                 mult=None, threads=1, verbose=False):
        L = [int(v) for v in (L if hasattr(L, "__len__") else [L])]
        M = [int(v) for v in (M if hasattr(M, "__len__") else [M])]
        dim = len(L)
        if mult is None:
            mult = 2 if family == 1 else 1
        arr, larr = c_size_t * dim, c_long * dim
        m = arr(*([0] * dim if m is None else [int(v) for v in m]))
        D = arr(*([0] * dim if D is None else [int(v) for v in D]))
        I = larr(*([-1] * dim if I is None else [int(v) for v in I]))
        self.dim, self.A, self.B = dim, A, B
        self._h = lib().ref_conv_create(dim, family, arr(*L), arr(*M), m, D, I, Sx, Sy,
                                        A, B, mult, threads, 1 if verbose else 0)
        self.doubles = int(lib().ref_conv_doubles(self._h))

    def close(self):
        if self._h:
            lib().ref_conv_destroy(self._h)
            self._h = None

    def params(self, d):
        buf = (c_size_t * 8)()
        lib().ref_conv_params(self._h, d, buf)
        return dict(zip("m p q n D inplace C S".split(), [int(v) for v in buf]))

    def convolve(self, arrays, normalized=True):
        n = max(self.A, self.B)
        ptrs = (c_void_p * n)(*[a.ctypes.data for a in arrays[:n]])
        lib().ref_conv_convolve(self._h, ptrs, 1 if normalized else 0)
        return arrays[0]

    def time(self, count):
        buf = (ctypes.c_double * count)()
        lib().ref_conv_time(self._h, count, buf)
        return [float(v) for v in buf]


def _dims(L):
    L = [int(v) for v in (L if hasattr(L, "__len__") else [L])]
    return len(L), (c_size_t * len(L))(*L)


def direct_complex(f, g):
    dim, L = _dims(f.shape)
    h = np.zeros_like(f)
    lib().ref_direct_complex(dim, L, f.ctypes.data, g.ctypes.data, h.ctypes.data)
    return h


def direct_real(f, g):
    dim, L = _dims(f.shape)
    h = np.zeros_like(f)
    lib().ref_direct_real(dim, L, f.ctypes.data, g.ctypes.data, h.ctypes.data)
    return h


def direct_centered1(f, g):
    h = np.zeros_like(f)
    lib().ref_direct_centered1(f.shape[0], f.ctypes.data, g.ctypes.data, h.ctypes.data)
    return h


def direct_hermitian(L, f, g):
    """L: logical lengths; f,g: symmetrised arrays shaped (Lx[,Ly],H_last)."""
    dim, Lc = _dims(L)
    h = np.zeros_like(f)
    lib().ref_direct_hermitian(dim, Lc, f.ctypes.data, g.ctypes.data, h.ctypes.data)
    return h


def symmetrize(L, f):
    dim, Lc = _dims(L)
    lib().ref_symmetrize(dim, Lc, f.ctypes.data)
    return f
