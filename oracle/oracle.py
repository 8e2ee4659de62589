"""ORACLE (test infrastructure; see oracle/__init__.py): numpy float64
restatement of what the hybrid dealiased convolution path computes.

Nothing here is imported by the product.  Every function cites the reference
file:line it follows (paths relative to /root/reference).

Parity pin: checked in tests/test_oracle.py against (a) the golden vectors in
tests/golden/ produced by running the reference itself in the build container
(tests/golden/make_golden.py), (b) the reference's own FFT-free direct
convolutions, and (c) the closed form of tests/hybridconv.cc:50-69.
"""
import ctypes
import os

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))

KIND_COMPLEX, KIND_CENTERED, KIND_HERMITIAN, KIND_REAL = 0, 1, 2, 3


def ceilquotient(a, b):
    return (a + b - 1) // b


# --------------------------------------------------------------------------
# parameter algebra
# --------------------------------------------------------------------------

def parameters(L, M, m, centered):
    """p, n, q of a padded FFT (convolve.cc:403-410)."""
    p = ceilquotient(L, m)
    P = p // 2 if ((centered and p % 2 == 0) or p == 2) else p
    n = ceilquotient(M, P * m)
    return p, n, P * n


def nextfftsize(m):
    """Smallest 2^a 3^b 5^c 7^d >= m (convolve.cc:114-124)."""
    def ceilpow2(x):
        v = 1
        while v < x:
            v *= 2
        return v
    N = ceilpow2(m)
    if m == N:
        return m
    a = 1
    while a < N:
        b = a
        while b < N:
            c = b
            while c < N:
                N = min(N, c * ceilpow2(ceilquotient(m, c)))
                c *= 3
            b *= 5
        a *= 7
    return N


def index_complex(r, i, *, m, p, q, n, D, D0, centered):
    """fftBase::index (convolve.h:297-326) with overwrite == false."""
    if q == 1:
        return i
    P = ceilquotient(p, 2)
    s = i % m
    if D > 1 and ((centered and p % 2 == 0) or p <= 2):
        u = (i // m) % P
        offset = 1 if (r == 0 and i >= P * m and D0 % 2 == 1) else 0
        incr = (i + P * m * offset) // (2 * P * m)
        r += incr
        if i // (P * m) - 2 * incr + offset == 1:
            if (not centered and p == 2) or (r > 0 and u == 0):
                s = s - 1 if s > 0 else m - 1
            if r == 0:
                r = n // 2
            else:
                r = n - r
                u = u - 1 if u > 0 else P - 1
    else:
        u = (i // m) % p
        r += i // (p * m)
    return q * s + n * u + r


def index_real(r, i, *, m, p, q, n):
    """fftPadReal::index (convolve.h:956-979), including the inner (p > 2)
    layouts."""
    if q == 1:
        return i
    s = i % m
    P = 1 if p == 2 else p
    r += i // (P * m)
    if p <= 2:
        if r == 0:
            return q * i
        if 2 * r == q:
            return q * m - (q * 2 * i + r)
    else:
        u = (i // m) % p
        if r == 0:
            if 2 * u == p:
                return q * m - (q * 2 * s + u * n)
            return u * n if s == 0 else q * m - (q * s - u * n)
        if 2 * r == n:
            return q * m - (q * s + 2 * u * n + r)
        return q * (m - s) - (u * n + r)
    return q * (m - s) - r


def real_blocksize(r, *, m, p, n):
    """fftPadReal::blocksize (convolve.h:941-945)."""
    e = m // 2 + 1
    if r == 0:
        if p > 2:
            return (p // 2 + 1) * m if p % 2 else (p // 2) * m + e - 1
        return e
    if 2 * r == n:
        return (p // 2) * m if p > 2 else e - 1
    return m * (1 if p == 2 else p)


# --------------------------------------------------------------------------
# explicit padded transforms (what every residue is compared against,
# tests/hybrid.cc:72-103, hybridh.cc, hybridr.cc:72-100)
# --------------------------------------------------------------------------

def padded_dft(kind, L, N, f):
    """Explicit padded DFT over the LAST axis... f has shape (Lin, C).

    complex:   F_k = sum_j f_j zeta_N^{+jk}                (convolve.cc:576: sign +1)
    centered:  F_k = sum_j f[j+floor(L/2)] zeta_N^{+jk}    (convolve.cc:2006-2036)
    hermitian: real F_k = sum_{|j|<H} g_j zeta_N^{+jk}, g_{-j}=conj(f_j)
               (convolve.cc:4439-4456: c2r)
    real:      F_k = sum_j x_j zeta_N^{-jk}, k <= N/2      (convolve.cc:5726-5743: r2c)
    Returns an array of shape (N or N//2+1, C).
    """
    f = np.asarray(f)
    C = f.shape[1]
    if kind == KIND_COMPLEX:
        pad = np.zeros((N, C), dtype=np.complex128)
        pad[:L] = f
        return np.fft.ifft(pad, axis=0) * N
    if kind == KIND_CENTERED:
        H = L // 2
        pad = np.zeros((N, C), dtype=np.complex128)
        for j in range(-H, L - H):
            pad[j % N] += f[j + H]
        return np.fft.ifft(pad, axis=0) * N
    if kind == KIND_HERMITIAN:
        H = ceilquotient(L, 2)
        pad = np.zeros((N, C), dtype=np.complex128)
        pad[0] = f[0].real
        for j in range(1, H):
            pad[j % N] += f[j]
            pad[(-j) % N] += np.conj(f[j])
        return (np.fft.ifft(pad, axis=0) * N).real
    pad = np.zeros((N, C), dtype=np.float64)
    pad[:L] = f
    return np.fft.rfft(pad, axis=0)


def real_spectrum_at(F2, N, i):
    """Value of the sign(-1) transform at index i from its half spectrum
    (tests/hybridr.cc:88-90)."""
    H = F2.shape[0]
    return F2[i] if i < H else np.conj(F2[N - i])


# --------------------------------------------------------------------------
# convolutions (definitions: tests/direct.h, tests/direct.cc)
# --------------------------------------------------------------------------

class _FFT:
    """numpy.fft, or scipy.fft (the same pocketfft algorithms, float64) run on
    all host cores when scipy is importable -- only the wall time of the
    full-size checks differs."""

    def __init__(self):
        try:
            import scipy.fft as sf
            try:
                w = max(1, len(os.sched_getaffinity(0)))
            except AttributeError:
                w = os.cpu_count() or 1
            self.kw = {"workers": w}
            self.m = sf
        except ImportError:
            self.kw = {}
            self.m = np.fft

    def fftn(self, a, s=None, axes=None):
        return self.m.fftn(a, s, axes=axes, **self.kw)

    def ifftn(self, a, s=None, axes=None):
        return self.m.ifftn(a, s, axes=axes, **self.kw)

    def rfftn(self, a, s=None, axes=None):
        return self.m.rfftn(a, s, axes=axes, **self.kw)

    def irfftn(self, a, s=None, axes=None):
        return self.m.irfftn(a, s, axes=axes, **self.kw)


_fft = _FFT()


def conv_complex(f, g):
    """h_i = sum_{j<=i} f_j g_{i-j} in every dimension (direct.h:19-26,77-88,
    122-136), evaluated with explicitly zero-padded FFTs."""
    f = np.asarray(f, dtype=np.complex128)
    g = np.asarray(g, dtype=np.complex128)
    shape = [2 * n - 1 for n in f.shape]
    ax = list(range(f.ndim))
    F = _fft.fftn(f, shape, axes=ax)
    F *= _fft.fftn(g, shape, axes=ax)
    h = _fft.ifftn(F, axes=ax)
    return np.ascontiguousarray(h[tuple(slice(0, n) for n in f.shape)])


def correlation_complex(f, g):
    """multcorrelation (convolve.cc:87-110): spectra combine as F conj(G)."""
    f = np.asarray(f, dtype=np.complex128)
    g = np.asarray(g, dtype=np.complex128)
    shape = [2 * n for n in f.shape]
    F = np.fft.ifftn(np.pad(f, [(0, s - n) for s, n in zip(shape, f.shape)]))
    G = np.fft.ifftn(np.pad(g, [(0, s - n) for s, n in zip(shape, g.shape)]))
    h = np.fft.fftn(F * np.conj(G)) * np.prod(shape)
    return np.ascontiguousarray(h[tuple(slice(0, n) for n in f.shape)])


def conv_real(f, g):
    f = np.asarray(f, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    shape = [2 * n - 1 for n in f.shape]
    ax = list(range(f.ndim))
    F = _fft.rfftn(f, shape, axes=ax)
    F *= _fft.rfftn(g, shape, axes=ax)
    h = _fft.irfftn(F, shape, axes=ax)
    return np.ascontiguousarray(h[tuple(slice(0, n) for n in f.shape)])


def conv_real_pencils(f, g, points):
    """z-pencils h[i,j,:] of the 3-D real linear convolution (direct.h:107-137)
    for the given (i,j): FFT along z only, direct sums over i' <= i, j' <= j.
    Lets a 512^3 result be checked in seconds without the full 3-D oracle."""
    f = np.asarray(f, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    Lz = f.shape[2]
    F = _fft.m.rfft(f, n=2 * Lz, axis=2, **_fft.kw)
    G = _fft.m.rfft(g, n=2 * Lz, axis=2, **_fft.kw)
    out = {}
    for (i, j) in points:
        H = np.einsum("abk,abk->k", F[:i + 1, :j + 1], G[i::-1, j::-1][:i + 1, :j + 1])
        out[(i, j)] = _fft.m.irfft(H, n=2 * Lz)[:Lz]
    return out


def _full_conv(a, b):
    shape = [x + y - 1 for x, y in zip(a.shape, b.shape)]
    ax = list(range(a.ndim))
    A = _fft.fftn(a, shape, axes=ax)
    A *= _fft.fftn(b, shape, axes=ax)
    return _fft.ifftn(A, axes=ax)


def conv_centered1(f, g):
    """Centred 1-D convolution (direct.h:27-39): index c <-> mode c-floor(L/2)."""
    L = f.shape[0]
    H = L // 2
    full = _full_conv(np.asarray(f, np.complex128), np.asarray(g, np.complex128))
    return np.ascontiguousarray(full[H:H + L])


def hermitian_full(L, f):
    """Expand an array holding the non-negative modes of its last axis
    (shape (Lx[,Ly],H)) to all modes |k| < H of that axis using
    g(-k) = conj(g(k)) over ALL axes (direct.cc:30-35,68-76)."""
    f = np.asarray(f, dtype=np.complex128)
    H = f.shape[-1]
    full = np.zeros(f.shape[:-1] + (2 * H - 1,), dtype=np.complex128)
    full[..., H - 1:] = f
    neg = np.conj(f[..., 1:][..., ::-1])  # modes -(H-1)..-1 of the last axis
    # reflect the centred leading axes about their origins
    for ax in range(f.ndim - 1):
        n = f.shape[ax]
        o = n // 2
        src = 2 * o - np.arange(n)
        ok = (src >= 0) & (src < n)
        taken = np.take(neg, np.clip(src, 0, n - 1), axis=ax)
        shape = [1] * neg.ndim
        shape[ax] = n
        neg = taken * ok.reshape(shape)
    full[..., :H - 1] = neg
    return full


def conv_hermitian(L, f, g):
    """Centred Hermitian convolution (direct.cc:5-98): leading axes centred at
    floor(L/2), last axis holds modes 0..H-1 of a real field."""
    L = [int(v) for v in (L if hasattr(L, "__len__") else [L])]
    ff = hermitian_full(L, f)
    gg = hermitian_full(L, g)
    full = _full_conv(ff, gg)
    H = f.shape[-1]
    sl = []
    for ax in range(f.ndim - 1):
        o = f.shape[ax] // 2
        sl.append(slice(o, o + f.shape[ax]))
    sl.append(slice(2 * (H - 1), 2 * (H - 1) + H))
    return np.ascontiguousarray(full[tuple(sl)])


def symmetrize(L, f):
    """HermitianSymmetrize / X / XY on an array shaped (Lx[,Ly],H)
    (convolve.h:1168-1267; origins floor(L/2) as in tests/hybridconvh*.cc)."""
    L = [int(v) for v in (L if hasattr(L, "__len__") else [L])]
    dim = len(L)
    if dim == 1:
        f[0] = f[0].real
        return f
    if dim == 2:
        Lx = L[0]
        Hx = ceilquotient(Lx, 2)
        x0 = Lx // 2
        for i in range(1, Hx):
            f[x0 - i, 0] = np.conj(f[x0 + i, 0])
        f[x0, 0] = f[x0, 0].real
        if x0 == Hx:
            f[0, :] = 0
        return f
    Lx, Ly = L[0], L[1]
    Hx, Hy = ceilquotient(Lx, 2), ceilquotient(Ly, 2)
    x0, y0 = Lx // 2, Ly // 2
    for i in range(1, Hx):
        f[x0 - i, y0, 0] = np.conj(f[x0 + i, y0, 0])
    f[x0, y0, 0] = f[x0, y0, 0].real
    for i in range(-Hx + 1, Hx):
        for j in range(1, Hy):
            f[x0 - i, y0 - j, 0] = np.conj(f[x0 + i, y0 + j, 0])
    if x0 == Hx:
        f[0, :, :] = 0
    if y0 == Hy:
        f[:, 0, :] = 0
    return f


def closed_form_1d(L):
    """Inputs and exact result of `hybridconv -a` (tests/hybridconv.cc:13-14,
    50-55,65-69): f=iF e^{ij}, g=iG e^{ij} => h_j=iF iG (j+1) e^{ij}."""
    iF = complex(np.sqrt(3.0), np.sqrt(7.0))
    iG = complex(np.sqrt(5.0), np.sqrt(11.0))
    j = np.arange(L)
    e = np.exp(1j * j)
    return iF * e, iG * e, iF * iG * (j + 1) * e


def rel_l2(a, b):
    """Relative L2 error as printed by the reference tests
    (tests/hybridconv.cc:118-129)."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.sqrt(np.sum(np.abs(b) ** 2))
    num = np.sqrt(np.sum(np.abs(a - b) ** 2))
    return float(num / den) if den > 0 else float(num)


def tolerance(*padded):
    """north_star tolerance: relative L2 <= 1e-12*log2(N), N = prod of padded sizes."""
    N = 1
    for v in padded:
        N *= int(v)
    return 1e-12 * max(1.0, np.log2(N))


# --------------------------------------------------------------------------
# C restatement of the direct sums (oracle/direct_oracle.c)
# --------------------------------------------------------------------------

_clib = None


def _c():
    global _clib
    if _clib is None:
        path = os.path.join(_here, "liboracle.so")
        if not os.path.exists(path):
            raise ImportError("oracle/liboracle.so is not built (make -C oracle oracle)")
        _clib = ctypes.CDLL(path)
    return _clib


def direct(kind, f, g, L=None):
    """FFT-free direct convolution via oracle/direct_oracle.c.
    kind: 'complex', 'real', 'centered' (1-D), 'hermitian' (L = logical lengths)."""
    f = np.ascontiguousarray(f)
    g = np.ascontiguousarray(g)
    h = np.zeros_like(f)
    dim = f.ndim
    sz = ctypes.c_size_t
    vp = ctypes.c_void_p
    args = [vp(f.ctypes.data), vp(g.ctypes.data), vp(h.ctypes.data)]
    if kind == "centered":
        _c().oracle_direct1_centered(sz(f.shape[0]), *args)
    elif kind == "hermitian":
        L = [int(v) for v in (L if hasattr(L, "__len__") else [L])]
        if dim == 1:
            _c().oracle_direct1_hermitian(sz(f.shape[0]), *args)
        elif dim == 2:
            _c().oracle_direct2_hermitian(sz(L[0]), sz(L[1]), *args)
        else:
            _c().oracle_direct3_hermitian(sz(L[0]), sz(L[1]), sz(L[2]), *args)
    else:
        fn = getattr(_c(), "oracle_direct%d_%s" % (dim, kind))
        fn(*[sz(n) for n in f.shape], *args)
    return h
