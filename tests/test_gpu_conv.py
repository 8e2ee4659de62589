"""GPU parity of Convolution / Convolution2 / Convolution3 against the oracle
(numpy explicit-padding FFT convolution, itself pinned to the reference in
tests/test_oracle.py), through the C ABI with host buffers.
Mirrors the reference's `-E` accuracy runs (tests/hybridconv*.cc with the
deterministic ramps) and tests/tests.py:561-653 parameter sweeps."""
import itertools

import numpy as np
import pytest

import fftwpp_b200 as fp
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def check(fam, L, M, m=None, D=None, I=None, mult=None, seed=0):
    rng = np.random.default_rng(1234 + seed)
    L = list(L)
    if fam == fp.FAMILY_COMPLEX:
        f, g = crand(rng, *L), crand(rng, *L)
        want = O.correlation_complex(f, g) if mult == fp.MULT_CORRELATION else O.conv_complex(f, g)
    elif fam == fp.FAMILY_REAL:
        f, g = rng.uniform(-1, 1, L), rng.uniform(-1, 1, L)
        want = O.conv_real(f, g)
    else:
        shp = L[:-1] + [(L[-1] + 1) // 2]
        f, g = crand(rng, *shp), crand(rng, *shp)
        O.symmetrize(L, f)
        O.symmetrize(L, g)
        want = O.conv_hermitian(L, f, g)
    conv = fp.HybridConv(L, M, family=fam, m=m, D=D, I=I, mult=mult)
    a = [np.ascontiguousarray(f.copy()), np.ascontiguousarray(g.copy())]
    conv.convolve(a)
    padded = [conv.params(d)["m"] * conv.params(d)["q"] for d in range(len(L))]
    err = O.rel_l2(a[0], want)
    assert err < O.tolerance(*padded), (err, [conv.params(d) for d in range(len(L))])
    conv.close()
    return err


@pytest.mark.parametrize("L", [1, 2, 3, 5, 7, 8, 16, 31, 64, 100, 512, 1000])
def test_conv1d_complex_auto(L):
    check(fp.FAMILY_COMPLEX, [L], [2 * L], seed=L)


def _forced_1d(kind_family):
    out = []
    centered = kind_family == fp.FAMILY_HERMITIAN
    for L in (3, 5, 7, 8, 12, 16, 24):
        Ms = [2 * L, 5 * L // 2, 3 * L] if not centered else [(3 * L + 1) // 2, 2 * L, 3 * L]
        for M in Ms:
            for m in sorted(set([M, L + 1, L, (L + 1) // 2, max(2, L // 4)])):
                p, n, q = O.parameters(L, M, m, centered)
                if q * m < M:
                    continue
                if kind_family == fp.FAMILY_COMPLEX:
                    Ds = [1] + ([n] if n > 1 else []) + ([2] if n > 2 else [])
                    if q == 1:
                        Ds = [1]
                elif kind_family == fp.FAMILY_HERMITIAN:
                    if q == 1:
                        Ds = [1]
                    elif p % 2 == 0:
                        Ds = [2]  # p > 2: the inner routines (C == 1)
                    else:
                        continue
                else:
                    if not ((n % 2 == 1 or p % 2 == 0 or p <= 2) and (q % 2 == 1 or m % 2 == 0)):
                        continue
                    Ds = [1]
                    if (n - 1) // 2 > 1:
                        Ds.append((n - 1) // 2)
                for D in Ds:
                    out.append((L, M, m, D))
    return out


@pytest.mark.parametrize("L,M,m,D", _forced_1d(fp.FAMILY_COMPLEX))
def test_conv1d_complex_forced(L, M, m, D):
    check(fp.FAMILY_COMPLEX, [L], [M], m=[m], D=[D], I=[0], seed=L + M + m)


@pytest.mark.parametrize("L,M,m,D", _forced_1d(fp.FAMILY_HERMITIAN))
def test_conv1d_hermitian_forced(L, M, m, D):
    check(fp.FAMILY_HERMITIAN, [L], [M], m=[m], D=[D], I=[0], seed=L + M + m)


@pytest.mark.parametrize("L,M,m,D", _forced_1d(fp.FAMILY_REAL))
def test_conv1d_real_forced(L, M, m, D):
    check(fp.FAMILY_REAL, [L], [M], m=[m], D=[D], I=[0], seed=L + M + m)


@pytest.mark.parametrize("L", [4, 7, 8, 33, 256])
def test_conv1d_hermitian_auto(L):
    check(fp.FAMILY_HERMITIAN, [L], None, seed=L)


@pytest.mark.parametrize("L", [4, 7, 8, 33, 256, 1024])
def test_conv1d_real_auto(L):
    check(fp.FAMILY_REAL, [L], [2 * L], seed=L)


def test_conv1d_correlation():
    check(fp.FAMILY_COMPLEX, [24], [48], mult=fp.MULT_CORRELATION)


@pytest.mark.parametrize("fam", [fp.FAMILY_COMPLEX, fp.FAMILY_REAL, fp.FAMILY_HERMITIAN])
@pytest.mark.parametrize("L", [(4, 4), (5, 7), (8, 6), (16, 16), (33, 20), (64, 128)])
def test_conv2d(fam, L):
    M = None if fam == fp.FAMILY_HERMITIAN else [2 * l for l in L]
    check(fam, L, M, seed=sum(L))


@pytest.mark.parametrize("fam", [fp.FAMILY_COMPLEX, fp.FAMILY_REAL, fp.FAMILY_HERMITIAN])
@pytest.mark.parametrize("L", [(4, 4, 4), (3, 5, 7), (8, 6, 5), (16, 16, 16), (12, 9, 20)])
def test_conv3d(fam, L):
    M = None if fam == fp.FAMILY_HERMITIAN else [2 * l for l in L]
    check(fam, L, M, seed=sum(L))


@pytest.mark.parametrize("fam", [fp.FAMILY_COMPLEX, fp.FAMILY_REAL])
def test_conv3d_forced_hybrid(fam):
    # explicit m per dimension: p=1 / p=2 / explicit mixes
    check(fam, (8, 8, 8), (16, 16, 16), m=[8, 4, 16], D=[1, 1, 1], I=[0, 0, 0])
    check(fam, (6, 10, 12), (12, 20, 24), m=[6, 5, 12], D=[1, 1, 1], I=[0, 0, 0])


def test_reference_ramp_3d_real():
    # the reference's own deterministic input (tests/hybridconvr3.cc:54-64)
    Lx = Ly = Lz = 4
    i, j, k = np.meshgrid(np.arange(Lx), np.arange(Ly), np.arange(Lz), indexing="ij")
    f = [(i + (a + 1) * j + a * k + 1).astype(np.float64) for a in range(2)]
    want = O.direct("real", f[0], f[1])
    conv = fp.HybridConv([Lx, Ly, Lz], [8, 8, 8], family=fp.FAMILY_REAL)
    a = [np.ascontiguousarray(f[0].copy()), np.ascontiguousarray(f[1].copy())]
    conv.convolve(a)
    assert O.rel_l2(a[0], want) < 1e-12


def test_wrapper_api_matches_generic():
    # reference wrapper entry points (wrappers/cfftw++.cc:54-66)
    import ctypes
    L = 37
    rng = np.random.default_rng(5)
    f, g = crand(rng, L), crand(rng, L)
    want = O.conv_complex(f, g)
    h = fp.lib.fftwpp_create_conv1d(L)
    a, b = f.copy(), g.copy()
    fp.lib.fftwpp_conv1d_convolve(h, a.ctypes.data, b.ctypes.data)
    fp.lib.fftwpp_conv1d_delete(h)
    assert O.rel_l2(a, want) < 1e-12


@pytest.mark.parametrize("L,m", [(256, 16), (1024, 32), (8192, 512), (8192, 128), (65536, 4096)])
@pytest.mark.parametrize("mult", [2, 3])
def test_conv1d_two_stage_inner(L, m, mult):
    """p > 2 with power-of-two m and p: the two-stage pipeline (reference
    forwardInner/backwardInner, convolve.cc:1227-1466,1765-1965)."""
    check(fp.FAMILY_COMPLEX, [L], [mult * L], m=[m], D=[1], I=[0], seed=L + m)


def test_cfg1_closed_form_full_size():
    """BASELINE configs[0]: L=2^20, M=2^21 against the exact solution of
    `hybridconv -a` (tests/hybridconv.cc:50-55,65-69)."""
    L = 1 << 20
    f, g, h = O.closed_form_1d(L)
    conv = fp.HybridConv([L], [2 * L])
    a = [f.copy(), g.copy()]
    conv.convolve(a)
    assert O.rel_l2(a[0], h) < O.tolerance(2 * L)
    # and against the numpy oracle on seeded random data
    rng = np.random.default_rng(1234)
    f, g = crand(rng, L), crand(rng, L)
    a = [f.copy(), g.copy()]
    conv.convolve(a)
    assert O.rel_l2(a[0], O.conv_complex(f, g)) < O.tolerance(2 * L)


def test_linearity_large_2d():
    """Size-independent property at a size the direct sums cannot reach:
    conv(a f1 + b f2, g) = a conv(f1,g) + b conv(f2,g)."""
    shape = (1024, 1024)
    rng = np.random.default_rng(99)
    f1, f2, g = crand(rng, *shape), crand(rng, *shape), crand(rng, *shape)
    conv = fp.HybridConv(list(shape), [2 * s for s in shape])
    outs = []
    for f in (f1, f2, 2.0 * f1 - 0.5j * f2):
        a = [np.ascontiguousarray(f.copy()), g.copy()]
        conv.convolve(a)
        outs.append(a[0])
    want = 2.0 * outs[0] - 0.5j * outs[1]
    assert O.rel_l2(outs[2], want) < 1e-13
    assert O.rel_l2(outs[0], O.conv_complex(f1, g)) < O.tolerance(2048, 2048)


@pytest.mark.parametrize("L,M,m", [(32, 48, 16), (256, 384, 128), (256, 512, 128), (1024, 1536, 512),
                                   (64, 128, 128), (255, 384, 128), (8190, 12288, 4096)])
def test_conv1d_hermitian_power_of_two(L, M, m):
    """Hermitian rows on the register kernels (two inputs share one complex
    FFT; reference forward2/backward2 + realMultBinary, convolve.cc:4517-4843)."""
    D = 1 if m >= M else 2
    check(fp.FAMILY_HERMITIAN, [L], [M], m=[m], D=[D], I=[0], seed=L + m)


def test_cfg3_shape_hermitian_3d():
    """BASELINE configs[2] geometry (centred x,y + Hermitian z, M=3L/2) at 64^3."""
    check(fp.FAMILY_HERMITIAN, (64, 64, 64), None, seed=64)
    check(fp.FAMILY_HERMITIAN, (128, 32, 256), None, seed=65)


def _strided(a, shape_alloc):
    """Embed array a in a larger zero array (the reference's -Sx/-Sy stride
    variants, tests/tests.py:393-459) and return (buffer, view)."""
    buf = np.zeros(shape_alloc, dtype=a.dtype)
    buf[tuple(slice(0, n) for n in a.shape)] = a
    return buf


@pytest.mark.parametrize("fam", [fp.FAMILY_COMPLEX, fp.FAMILY_REAL])
@pytest.mark.parametrize("L", [(8, 6), (16, 16), (33, 12), (64, 32)])
def test_conv2d_with_x_stride(fam, L):
    Lx, Ly = L
    Sx = Ly + 3 if fam == fp.FAMILY_COMPLEX else Ly + 2
    rng = np.random.default_rng(Lx * Ly)
    if fam == fp.FAMILY_COMPLEX:
        f, g = crand(rng, Lx, Ly), crand(rng, Lx, Ly)
        want = O.conv_complex(f, g)
    else:
        f, g = rng.uniform(-1, 1, L), rng.uniform(-1, 1, L)
        want = O.conv_real(f, g)
    conv = fp.HybridConv([Lx, Ly], [2 * Lx, 2 * Ly], family=fam, Sx=Sx)
    a = [_strided(f, (Lx, Sx)), _strided(g, (Lx, Sx))]
    conv.convolve(a)
    assert O.rel_l2(a[0][:, :Ly], want) < 1e-12


@pytest.mark.parametrize("fam", [fp.FAMILY_COMPLEX, fp.FAMILY_REAL])
@pytest.mark.parametrize("L,gaps", [((6, 5, 4), (0, 3)), ((6, 5, 4), (2, 0)), ((6, 5, 4), (2, 3)),
                                    ((16, 16, 16), (1, 2)), ((32, 8, 64), (4, 4))])
def test_conv3d_with_strides(fam, L, gaps):
    """Sy > Lz exercises the non-contiguous x pass of Convolution3
    (reference convolve.h:1727-1735), Sx > Ly*Sy the x stride."""
    Lx, Ly, Lz = L
    Sy = Lz + gaps[0]
    Sx = Ly * Sy + gaps[1]
    if fam == fp.FAMILY_REAL:
        # doubles: keep strides even so Complex-typed offsets stay aligned
        Sy += Sy % 2
        Sx = Ly * Sy + 2 * (gaps[1] // 2)
    rng = np.random.default_rng(Lx + Ly + Lz)
    if fam == fp.FAMILY_COMPLEX:
        f, g = crand(rng, *L), crand(rng, *L)
        want = O.conv_complex(f, g)
    else:
        f, g = rng.uniform(-1, 1, L), rng.uniform(-1, 1, L)
        want = O.conv_real(f, g)
    conv = fp.HybridConv(list(L), [2 * l for l in L], family=fam, Sx=Sx, Sy=Sy)

    def embed(a):
        buf = np.zeros(Lx * Sx, dtype=a.dtype)
        for i in range(Lx):
            for j in range(Ly):
                buf[i * Sx + j * Sy: i * Sx + j * Sy + Lz] = a[i, j]
        return buf

    a = [embed(f), embed(g)]
    conv.convolve(a)
    got = np.array([[a[0][i * Sx + j * Sy: i * Sx + j * Sy + Lz] for j in range(Ly)]
                    for i in range(Lx)])
    assert O.rel_l2(got, want) < 1e-12


@pytest.mark.parametrize("A,B", [(1, 1), (3, 3), (3, 2)])
def test_multnone_general_A_B(A, B):
    """multNone with general A >= B: forward/backward identity times N
    (Application(A,B,multNone), reference convolve.h:84-121)."""
    L = 24
    rng = np.random.default_rng(A * 10 + B)
    arrays = [crand(rng, L) for _ in range(max(A, B))]
    keep = [a.copy() for a in arrays]
    conv = fp.HybridConv([L], [2 * L], A=A, B=B, mult=fp.MULT_NONE)
    conv.convolve(arrays)
    for b in range(B):
        assert O.rel_l2(arrays[b], keep[b]) < 1e-12


def test_async_pipeline_matches_blocking_call():
    """convolve_async/wait (two slots, pinned host buffers) returns the same
    results as the blocking call, step after step."""
    L = (32, 16, 24)
    rng = np.random.default_rng(77)
    conv = fp.HybridConv(list(L), [2 * l for l in L], family=fp.FAMILY_REAL)
    sets, wants = [], []
    for step in range(5):
        f, g = rng.uniform(-1, 1, L), rng.uniform(-1, 1, L)
        wants.append(O.conv_real(f, g))
        bufs = [fp.pinned_array(L, np.float64) for _ in range(2)]
        bufs[0][...] = f
        bufs[1][...] = g
        sets.append(bufs)
    for step in range(5):
        s = step % 2
        conv.wait(s)
        if step >= 2:
            assert O.rel_l2(sets[step - 2][0], wants[step - 2]) < 1e-12
        conv.convolve_async(sets[step], slot=s)
    conv.wait(0)
    conv.wait(1)
    for step in (3, 4):
        assert O.rel_l2(sets[step][0], wants[step]) < 1e-12
    conv.close()


@pytest.mark.parametrize("L,M,mult", [(512, 1024, None), (400, 800, None), (512, 1536, None),
                                      (300, 1024, fp.MULT_CORRELATION)])
def test_conv_rows_m512_batches(L, M, mult):
    """Batched rows on the m=512, p=1 fused row kernels (the z pass of cfg4):
    ragged row counts, L < m, q = 2 and 3, both built-in multipliers."""
    import torch
    rng = np.random.default_rng(L + M)
    conv = fp.HybridConv([L], [M], m=[512], D=[1], I=[0], mult=mult)
    for rows in (1, 7, 19, 300):
        f, g = crand(rng, rows, L), crand(rng, rows, L)
        td = [torch.from_numpy(f.copy()).cuda(), torch.from_numpy(g.copy()).cuda()]
        conv.convolve_rows(td, rows, L)
        torch.cuda.synchronize()
        got = td[0].cpu().numpy()
        for i in (0, rows // 2, rows - 1):
            want = (O.correlation_complex(f[i], g[i]) if mult == fp.MULT_CORRELATION
                    else O.conv_complex(f[i], g[i]))
            assert O.rel_l2(got[i], want) < 1e-12, (rows, i)
    conv.close()
