"""User multipliers (reference `multiplier` callbacks handed to Application,
convolve.h:78-121; advertised uses: autoconvolution / ternary products,
README.md:67): the host-callback path and the device-callback path must give
the same result as the fused built-ins and as the oracle."""
import ctypes

import numpy as np
import pytest
import torch

import fftwpp_b200 as fp
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def host_binary(F, n, r, offset):
    F[0][:] = F[0] * F[1]


def host_ternary(F, n, r, offset):
    F[0][:] = F[0] * F[1] * F[2]


class _DevView:
    """__cuda_array_interface__ view of n doubles at a raw device address."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 2}


def _dev(ptr, n, cplx=True):
    t = torch.as_tensor(_DevView(ptr, 2 * n if cplx else n), device="cuda")
    return torch.view_as_complex(t.view(n, 2)) if cplx else t


def device_binary(ptrs, n, r, offset, stream):
    a, b = _dev(ptrs[0], n), _dev(ptrs[1], n)
    a.mul_(b)


def device_real_binary(ptrs, n, r, offset, stream):
    a, b = _dev(ptrs[0], n, False), _dev(ptrs[1], n, False)
    a.mul_(b)


@pytest.mark.parametrize("L", [8, 33, 256])
@pytest.mark.parametrize("dev", [False, True])
def test_custom_binary_1d_complex(L, dev):
    rng = np.random.default_rng(L)
    f, g = crand(rng, L), crand(rng, L)
    want = O.conv_complex(f, g)
    conv = fp.HybridConv([L], [2 * L], mult=host_binary,
                         device_mult=device_binary if dev else None)
    a = [f.copy(), g.copy()]
    conv.convolve(a)
    assert O.rel_l2(a[0], want) < 1e-12
    conv.close()


@pytest.mark.parametrize("dev", [False, True])
def test_custom_binary_forced_residue_blocks(dev):
    """D > 1 and p > 2: several multiplier calls per residue pass."""
    L, M = 24, 60
    rng = np.random.default_rng(5)
    f, g = crand(rng, L), crand(rng, L)
    want = O.conv_complex(f, g)
    for m, D in ((8, 1), (4, 2), (24, 1)):
        conv = fp.HybridConv([L], [M], m=[m], D=[D], I=[0], mult=host_binary,
                             device_mult=device_binary if dev else None)
        a = [f.copy(), g.copy()]
        conv.convolve(a)
        assert O.rel_l2(a[0], want) < 1e-12, (m, D)
        conv.close()


def test_custom_ternary_A3():
    """A=3, B=1 with M >= 3L-2: dealiased product of three sequences."""
    L = 12
    rng = np.random.default_rng(3)
    f, g, h = crand(rng, L), crand(rng, L), crand(rng, L)
    full = np.convolve(np.convolve(f, g), h)[:L]
    conv = fp.HybridConv([L], [3 * L - 2], A=3, B=1, mult=host_ternary)
    a = [f.copy(), g.copy(), h.copy()]
    conv.convolve(a)
    assert O.rel_l2(a[0], full) < 1e-12
    conv.close()


@pytest.mark.parametrize("dev", [False, True])
def test_custom_real_and_hermitian(dev):
    rng = np.random.default_rng(11)
    L = 16
    f, g = rng.uniform(-1, 1, L), rng.uniform(-1, 1, L)
    conv = fp.HybridConv([L], [2 * L], family=fp.FAMILY_REAL, mult=host_binary,
                         device_mult=device_binary if dev else None)
    a = [f.copy(), g.copy()]
    conv.convolve(a)
    assert O.rel_l2(a[0], O.conv_real(f, g)) < 1e-12
    conv.close()
    # Hermitian: the multiplier sees REAL transformed data (realMultBinary)
    H = (L + 1) // 2
    f, g = crand(rng, H), crand(rng, H)
    O.symmetrize([L], f)
    O.symmetrize([L], g)

    def host_real(F, n, r, offset):
        F[0][:] = F[0] * F[1]

    conv = fp.HybridConv([L], None, family=fp.FAMILY_HERMITIAN, mult=host_real,
                         device_mult=device_real_binary if dev else None)
    a = [f.copy(), g.copy()]
    conv.convolve(a)
    assert O.rel_l2(a[0], O.conv_hermitian([L], f, g)) < 1e-12
    conv.close()


@pytest.mark.parametrize("dev", [False, True])
def test_custom_binary_2d_device_arrays(dev):
    """2-D with the custom multiplier on the innermost dimension; device
    tensors in, so with a device multiplier nothing touches the host."""
    Lx, Ly = 8, 12
    rng = np.random.default_rng(9)
    f, g = crand(rng, Lx, Ly), crand(rng, Lx, Ly)
    want = O.conv_complex(f, g)
    conv = fp.HybridConv([Lx, Ly], [2 * Lx, 2 * Ly], mult=host_binary,
                         device_mult=device_binary if dev else None)
    a = [torch.from_numpy(f).cuda(), torch.from_numpy(g).cuda()]
    conv.convolve(a)
    torch.cuda.synchronize()
    assert O.rel_l2(a[0].cpu().numpy(), want) < 1e-12
    conv.close()


# ---------------------------------------------------------------------------
# index-aware multipliers (reference Indices, convolve.h:48-76; the outer
# levels set indices.index[d] per transformed row, convolve.h:1442,1759; the
# innermost index is fft->index(r,j+offset), convolve.cc:39-48)
# ---------------------------------------------------------------------------

def _weight(*k):
    """Some function of the transformed multi-index (kx[,ky[,kz]])."""
    t = sum((d + 2) * np.asarray(kk) for d, kk in enumerate(k))
    return 1.0 / (1.0 + (t % 7)) + 0.25j * ((t % 5) - 2)


def _filtered_oracle(f, g, padded):
    """fft(F G W)/N restricted to [0,L): forward sign +1 (convolve.cc:576),
    F = explicit padded DFT, W evaluated on the padded index grid."""
    ax = tuple(range(f.ndim))
    N = int(np.prod(padded))
    F = np.fft.ifftn(f, padded, axes=ax) * N
    G = np.fft.ifftn(g, padded, axes=ax) * N
    grids = np.meshgrid(*[np.arange(n) for n in padded], indexing="ij")
    h = np.fft.fftn(F * G * _weight(*grids), axes=ax) / N
    return h[tuple(slice(0, n) for n in f.shape)]


def _indexed_host(F, n, ctx):
    kin = np.array([ctx.index(j) for j in range(n)])
    outer = ctx.outer[::-1]          # reference order: outermost (x) last
    F[0][:] = F[0] * F[1] * _weight(*outer, kin)


def _indexed_device(ptrs, n, ctx, stream):
    kin = np.array([ctx.index(j) for j in range(n)])
    w = torch.from_numpy(np.asarray(_weight(*ctx.outer[::-1], kin), dtype=np.complex128)).cuda()
    a, b = _dev(ptrs[0], n), _dev(ptrs[1], n)
    a.mul_(b).mul_(w)


@pytest.mark.parametrize("dev", [False, True])
@pytest.mark.parametrize("L,M,m,D", [((12,), (24,), None, None), ((12,), (30,), (4,), (1,)),
                                     ((8,), (32,), (8,), (2,)),
                                     ((6, 8), (12, 16), None, None), ((6, 8), (12, 20), (3, 4), (1, 1)),
                                     ((4, 5, 6), (8, 10, 12), None, None),
                                     ((4, 6, 8), (8, 12, 16), (2, 6, 4), (1, 1, 2))])
def test_index_aware_multiplier(L, M, m, D, dev):
    rng = np.random.default_rng(sum(L) + sum(M))
    f, g = crand(rng, *L), crand(rng, *L)
    conv = fp.HybridConv(list(L), list(M), m=m, D=D, I=None if m is None else [0] * len(L),
                         mult=_indexed_host, device_mult=_indexed_device if dev else None,
                         indexed=True)
    padded = [conv.params(d)["m"] * conv.params(d)["q"] for d in range(len(L))]
    want = _filtered_oracle(f, g, padded)
    a = [np.ascontiguousarray(f.copy()), np.ascontiguousarray(g.copy())]
    conv.convolve(a)
    assert O.rel_l2(a[0], want) < 1e-12, [conv.params(d) for d in range(len(L))]
    conv.close()


def test_many_arrays_beyond_fused_limit():
    """A=10 > 8 arrays: the built-in multNone runs on the unfused path instead
    of overflowing the fused kernels' argument block."""
    L, A = 16, 10
    rng = np.random.default_rng(3)
    arrays = [crand(rng, L) for _ in range(A)]
    keep = [a.copy() for a in arrays]
    conv = fp.HybridConv([L], [2 * L], A=A, B=A, mult=fp.MULT_NONE)
    conv.convolve(arrays)
    for a, k in zip(arrays, keep):
        assert O.rel_l2(a, k) < 1e-12
    conv.close()
