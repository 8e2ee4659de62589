"""The PRODUCT's HermitianSymmetrize / X / XY (cpp/convolve.cc; reference
convolve.h:1168-1267) through the C ABI of lib_fftwpp.so, against the numpy
restatement and, when oracle/_ref is built, the reference's own routine.
Host-side code: runs without a GPU."""
import ctypes

import numpy as np
import pytest

import fftwpp_b200 as fp
from oracle import oracle as O
from oracle import ref as R

SHAPES_2D = [(4, 3), (5, 4), (7, 7), (8, 1), (2, 5), (16, 9), (33, 6)]
SHAPES_3D = [(4, 4, 3), (5, 7, 2), (7, 4, 4), (8, 8, 1), (2, 2, 5), (9, 16, 5), (12, 5, 7)]


def crand(rng, shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def product_symmetrize(a):
    """Call the exported wrapper exactly as wrappers/fftwpp.py:327-349 does."""
    lib = fp.lib
    p = ctypes.c_void_p(a.ctypes.data)
    if a.ndim == 1:
        lib.fftwpp_HermitianSymmetrize(p)
    elif a.ndim == 2:
        Lx, Hy = a.shape
        lib.fftwpp_HermitianSymmetrizeX((Lx + 1) // 2, Hy, Lx // 2, p)
    else:
        Lx, Ly, Hz = a.shape
        lib.fftwpp_HermitianSymmetrizeXY((Lx + 1) // 2, (Ly + 1) // 2, Hz, Lx // 2, Ly // 2, p)
    return a


def logical(shape):
    # last axis holds modes 0..H-1 of L = 2H-1 (odd) logical modes
    return list(shape[:-1]) + [2 * shape[-1] - 1]


@pytest.mark.parametrize("shape", [(1,), (5,)] + SHAPES_2D + SHAPES_3D)
def test_product_symmetrize_matches_oracle(shape):
    rng = np.random.default_rng(sum(shape))
    a = np.ascontiguousarray(crand(rng, shape))
    want = O.symmetrize(logical(shape), a.copy())
    got = product_symmetrize(a.copy())
    assert np.array_equal(got, want)
    # idempotent, and untouched outside the symmetry plane
    assert np.array_equal(product_symmetrize(got.copy()), got)
    if len(shape) == 3 and shape[2] > 1:
        keep = np.ones(shape, bool)
        keep[:, :, 0] = False
        if shape[0] % 2 == 0:
            keep[0] = False
        if shape[1] % 2 == 0:
            keep[:, 0] = False
        assert np.array_equal(got[keep], a[keep])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", [(6,)] + SHAPES_2D + SHAPES_3D)
def test_product_symmetrize_matches_reference(shape):
    rng = np.random.default_rng(100 + sum(shape))
    a = np.ascontiguousarray(crand(rng, shape))
    want = R.symmetrize(logical(shape), a.copy())
    got = product_symmetrize(a.copy())
    assert np.array_equal(got, want)


def test_symmetrized_field_is_real():
    """Property: after symmetrisation the centred half-spectrum describes a
    real field (the reason the reference applies it before hconv)."""
    shape = (7, 9, 4)
    rng = np.random.default_rng(3)
    a = product_symmetrize(np.ascontiguousarray(crand(rng, shape)))
    full = O.hermitian_full(logical(shape), a)
    spec = np.fft.ifftshift(full, axes=(0, 1, 2))
    field = np.fft.ifftn(spec)
    assert np.max(np.abs(field.imag)) < 1e-14 * np.max(np.abs(field.real))
