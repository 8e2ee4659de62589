"""All six wrapper families of the reference C API (reference
wrappers/cfftw++.cc:54-163) and the Python binding module that mirrors
reference wrappers/fftwpp.py, with the array shapes of its doctests
(fftwpp.py:100-199,205-320), against the oracle."""
import numpy as np
import pytest

import fftwpp_b200 as fp
from fftwpp_b200 import fftwpp as W
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def fill(shape, seed):
    rng = np.random.default_rng(seed)
    a = W.complex_align(shape)
    a[...] = rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)
    return a


@pytest.mark.parametrize("shape", [8, (8,), (37,), (4, 4), (6, 9), (4, 4, 4), (5, 3, 8), (32, 32, 32)])
def test_convolution_wrapper(shape):
    shp = (shape,) if isinstance(shape, int) else shape
    f, g = fill(shp, 1), fill(shp, 2)
    assert f.ctypes.data % 16 == 0
    want = O.conv_complex(f, g)
    c = W.Convolution(shape)
    c.convolve(f, g)
    assert O.rel_l2(f, want) < 1e-12
    del c


@pytest.mark.parametrize("shape", [(4,), (9,), (7, 4), (8, 4), (7, 7, 4), (8, 6, 5), (15, 16, 8)])
def test_hconvolution_wrapper(shape):
    """Hermitian wrappers: arrays (Lx[,Ly],Hz) <-> logical last length 2Hz-1;
    the wrapper symmetrises its inputs itself (fftwpp.py:327-349)."""
    f, g = fill(shape, 3), fill(shape, 4)
    L = list(shape[:-1]) + [2 * shape[-1] - 1]
    fs, gs = O.symmetrize(L, f.copy()), O.symmetrize(L, g.copy())
    want = O.conv_hermitian(L, fs, gs)
    c = W.HConvolution(shape)
    c.convolve(f, g)
    assert O.rel_l2(f, want) < 1e-12
    del c


def test_reference_doctest_inputs():
    """The deterministic inputs of the reference's doctests
    (fftwpp.py:112-117,160-166)."""
    N = 8
    f, g = W.complex_align([N]), W.complex_align([N])
    for i in range(N):
        f[i] = complex(i + 1, i + 3)
        g[i] = complex(i + 2, 2 * i + 3)
    want = O.direct("complex", f.copy(), g.copy())
    W.Convolution(N).convolve(f, g)
    assert O.rel_l2(f, want) < 1e-13
    N = 4
    f, g = W.complex_align([N, N, N]), W.complex_align([N, N, N])
    i, j, k = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    f[...] = (i + 1) + 1j * (j + 3 + k)
    g[...] = (i + k + 1) + 1j * (2 * j + 3 + k)
    want = O.direct("complex", f.copy(), g.copy())
    W.Convolution(f.shape).convolve(f, g)
    assert O.rel_l2(f, want) < 1e-13


def test_maxthreads_roundtrip():
    old = W.fftwpp_get_maxthreads()
    W.fftwpp_set_maxthreads(3)
    assert W.fftwpp_get_maxthreads() == 3
    W.fftwpp_set_maxthreads(old)


@pytest.mark.parametrize("name,dims", [("conv1d", (21,)), ("hconv1d", (21,)), ("conv2d", (6, 10)),
                                       ("hconv2d", (7, 9)), ("conv3d", (4, 6, 5)),
                                       ("hconv3d", (5, 7, 9))])
def test_raw_c_entry_points(name, dims):
    """create / convolve / delete of every family straight through ctypes, as
    wrappers/cexample.c drives them."""
    import ctypes
    herm = name.startswith("h")
    shape = tuple(dims[:-1]) + (((dims[-1] + 1) // 2) if herm else dims[-1],)
    f, g = fill(shape, 11), fill(shape, 12)
    if herm:
        O.symmetrize(list(dims), f)
        O.symmetrize(list(dims), g)
        want = O.conv_hermitian(list(dims), f, g)
    else:
        want = O.conv_complex(f, g)
    lib = fp.lib
    h = getattr(lib, "fftwpp_create_" + name)(*dims)
    assert h
    getattr(lib, "fftwpp_%s_convolve" % name)(h, ctypes.c_void_p(f.ctypes.data),
                                              ctypes.c_void_p(g.ctypes.data))
    getattr(lib, "fftwpp_%s_delete" % name)(h)
    assert O.rel_l2(f, want) < 1e-12
