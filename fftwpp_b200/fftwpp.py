"""Drop-in for the reference's Python binding module (reference
wrappers/fftwpp.py:14-25,100-351): the same public names -- complex_align,
Convolution, HConvolution, fftwpp_set_maxthreads, fftwpp_get_maxthreads --
bound with ctypes to the same C entry points (reference wrappers/cfftw++.cc:
27-163), which here live in the B200 library lib_fftwpp.so.

Shapes follow the reference: Convolution(shape) convolves complex arrays of
that shape in place into f; HConvolution(shape) takes centred Hermitian data
whose LAST axis holds only the non-negative modes, i.e. arrays shaped
(Lx[,Ly],Hz) describe Lz=2*Hz-1 modes (reference fftwpp.py:295-320), and
enforces the Hermitian symmetry of both inputs before convolving
(reference fftwpp.py:327-349).
"""
import ctypes

import numpy as np
from numpy.ctypeslib import ndpointer

from ._lib import lib_path

# a private handle: its prototypes (numpy ndpointer arguments, as in the
# reference module) do not disturb the package-wide ctypes signatures
clib = ctypes.CDLL(lib_path)
clib.get_fftwpp_maxthreads.restype = ctypes.c_size_t

__all__ = ["Convolution", "HConvolution", "complex_align",
           "fftwpp_set_maxthreads", "fftwpp_get_maxthreads"]

_c128 = ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
_sz = ctypes.c_size_t


def complex_align(shape):
    """complex128 array of the given shape whose data is 16-byte aligned."""
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    count = int(np.prod(shape))
    raw = np.empty(count * 16 + 16, dtype=np.uint8)
    skip = (-raw.ctypes.data) % 16
    return raw[skip:skip + 16 * count].view(np.complex128).reshape(shape)


def fftwpp_set_maxthreads(nthreads):
    clib.set_fftwpp_maxthreads(_sz(nthreads))


def fftwpp_get_maxthreads():
    return int(clib.get_fftwpp_maxthreads())


def _bind(prefix, dim):
    create = getattr(clib, "fftwpp_create_%s%dd" % (prefix, dim))
    conv = getattr(clib, "fftwpp_%s%dd_convolve" % (prefix, dim))
    delete = getattr(clib, "fftwpp_%s%dd_delete" % (prefix, dim))
    create.restype = ctypes.c_void_p
    create.argtypes = [_sz] * dim
    conv.restype = None
    conv.argtypes = [ctypes.c_void_p, _c128, _c128]
    delete.restype = None
    delete.argtypes = [ctypes.c_void_p]
    return create, conv, delete


class _Base(object):
    prefix = None

    def __init__(self, shape):
        shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
        if not 1 <= len(shape) <= 3:
            raise ValueError("invalid shape (length/dimension should be 1, 2, or 3)")
        self.dim = len(shape)
        self.shape = shape
        create, self._convolve, self._delete = _bind(self.prefix, self.dim)
        self.cptr = create(*self._lengths(shape))

    def _lengths(self, shape):
        return shape

    def __del__(self):
        cptr, self.cptr = getattr(self, "cptr", None), None
        if cptr:
            self._delete(cptr)

    def _check(self, f, g):
        if tuple(f.shape) != self.shape or tuple(g.shape) != self.shape:
            raise AssertionError("arrays must have shape %s" % (self.shape,))


class Convolution(_Base):
    """Implicitly zero-padded complex convolution; f is overwritten."""
    prefix = "conv"

    def convolve(self, f, g):
        self._check(f, g)
        self._convolve(self.cptr, f, g)


class HConvolution(_Base):
    """Implicitly zero-padded centred Hermitian-symmetric convolution."""
    prefix = "hconv"

    def _lengths(self, shape):
        # the last axis stores modes 0..H-1 of 2H-1 (reference fftwpp.py:297-314)
        return shape[:-1] + (2 * shape[-1] - 1,)

    def convolve(self, f, g):
        self._check(f, g)
        if self.dim == 1:
            sym = clib.fftwpp_HermitianSymmetrize
            sym.argtypes = [_c128]
            sym(f)
            sym(g)
        elif self.dim == 2:
            Lx, Hy = f.shape
            sym = clib.fftwpp_HermitianSymmetrizeX
            sym.argtypes = [_sz, _sz, _sz, _c128]
            for a in (f, g):
                sym((Lx + 1) // 2, Hy, Lx // 2, a)
        else:
            Lx, Ly, Hz = f.shape
            sym = clib.fftwpp_HermitianSymmetrizeXY
            sym.argtypes = [_sz, _sz, _sz, _sz, _sz, _c128]
            for a in (f, g):
                sym((Lx + 1) // 2, (Ly + 1) // 2, Hz, Lx // 2, Ly // 2, a)
        self._convolve(self.cptr, f, g)
