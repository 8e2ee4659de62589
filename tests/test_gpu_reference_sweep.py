"""The parameter sweep of the reference's own unit-test driver, restated.

tests/tests.py (reference) runs the programs hybrid, hybridh, hybridr (forward /
backward identity) and hybridconv{,2,3}, hybridconvh{,2,3}, hybridconvr{,2,3}
over fixed (L, M, m) lists per transform class (complexTests, centeredTests,
hermitianTests, realTests: tests.py:560-652), every admissible D
(collectTests/getDs, tests.py:459-536), I in {0, 1}, and -- with -S -- one extra
word of stride.  This file walks the same cases through the GPU path: the 1-D
lists in full detail, the multi-dimensional ones as tests.py does by default
(outermost dimension in detail, inner dimensions on their basic case)."""
import numpy as np
import pytest

import fftwpp_b200 as fp
from oracle import oracle as O
from test_gpu_conv import check
from test_gpu_pad import check_pad

pytestmark = pytest.mark.gpu

COMPLEX, CENTERED, HERMITIAN, REAL = "s", "c", "H", "r"
KIND = {COMPLEX: 0, CENTERED: 1, HERMITIAN: 2, REAL: 3}


def cq(a, b):
    return -(-a // b)


def lmm(ttype, det):
    """(L, M, m) lists of complexTests / centeredTests / hermitianTests / realTests."""
    out = []
    if ttype == COMPLEX:
        L = 8
        for M in [2 * L] + ([cq(5 * L, 2)] if det else []):
            ms = [M, L + 1, L, cq(L, 2), cq(L, 4)] if det else [L]
            out += [(L, M, m) for m in ms]
    elif ttype == CENTERED:
        for L in [8] + ([7] if det else []):
            L2 = cq(L, 2)
            M0 = 3 * L2 - 2 * (L % 2)
            for M in [M0] + ([M0 + 1, 2 * L, 5 * L2] if det else []):
                ms = [cq(L, 4), L2, cq(L2 + L, 2), M] if det else [L2]
                out += [(L, M, m) for m in ms]
    elif ttype == HERMITIAN:
        for L in [8] + ([7] if det else []):
            L2 = cq(L, 2)
            M0 = 3 * L2 - 2 * (L % 2)
            for M in [M0] + ([2 * L, 5 * L2] if det else []):
                ms = [M, L2, cq(L, 4)] if det else [L2]
                out += [(L, M, m) for m in ms]
    else:
        out.append((8, 16, 16))                       # explicit
        Ls, Ms = ([8, 3], [16, 24, 64]) if det else ([8], [16])
        out += [(L, M, 8) for M in Ms for L in Ls]    # p = 1
        if det:
            out += [(L, M, 4) for M in (16, 24, 32) for L in (8, 5)]          # p = 2
            out += [(L, 16, 2) for L in (8, 7)]                                # p > 2
            out += [(L, 63, 3) for L in (9, 7)]
            out += [(L, 96, 4) for L in (24, 21)]
    return out


def Ds(ttype, L, M, m, C, mult=True):
    """collectTests + getDs: the D values tests.py runs for one (L, M, m, C)."""
    centered = ttype in (CENTERED, HERMITIAN)
    p, n, q = O.parameters(L, M, m, centered)
    if q == 1:
        return [1]
    if C > 1:
        return [2 if (ttype == HERMITIAN and not mult) else 1]
    if ttype == HERMITIAN:
        return [2]
    stop = (n - 1) // 2 if ttype == REAL else n
    start = 1
    out = [start] + list(range(start + 2 - start % 2, stop, 2))
    if stop > start:
        out.append(stop)
    return out


def valid(ttype, L, M, m, C, S, D):
    centered = ttype in (CENTERED, HERMITIAN)
    p, n, q = O.parameters(L, M, m, centered)
    if q * m < M:
        return False
    if q == 1:
        return D == 1
    if ttype in (COMPLEX, CENTERED):
        ok = D == 1 or (S == 1 and ((D < n and D % 2 == 0) or D == n))
        return ok and (ttype == COMPLEX or p % 2 == 0)
    if ttype == HERMITIAN:
        return D == 2 and p % 2 == 0 and (p == 2 or C == 1)
    return ((n % 2 == 1 or p % 2 == 0 or p <= 2) and (q % 2 == 1 or m % 2 == 0)
            and (D == 1 or (S == 1 and ((D < (n - 1) // 2 and D % 2 == 0) or D == (n - 1) // 2))))


def cases_1d(ttype, C=1, S=1, mult=True):
    out = []
    for (L, M, m) in lmm(ttype, True):
        for D in Ds(ttype, L, M, m, C, mult):
            if valid(ttype, L, M, m, C, S, D):
                out.append((L, M, m, D))
    return sorted(set(out))


FAMILY = {COMPLEX: fp.FAMILY_COMPLEX, HERMITIAN: fp.FAMILY_HERMITIAN, REAL: fp.FAMILY_REAL}


@pytest.mark.parametrize("ttype", [COMPLEX, HERMITIAN, REAL])
def test_hybridconv_1d_sweep(ttype):
    """hybridconv / hybridconvh / hybridconvr (tests.py dim 1), I in {0,1}."""
    n = 0
    for (L, M, m, D) in cases_1d(ttype):
        for I in (0, 1):
            check(FAMILY[ttype], [L], [M], m=[m], D=[D], I=[I], seed=L + M + m + D)
            n += 1
    assert n >= 20


@pytest.mark.parametrize("ttype", [COMPLEX, CENTERED, HERMITIAN, REAL])
@pytest.mark.parametrize("C,S", [(1, 1), (2, 2), (2, 3)])
def test_hybrid_identity_sweep(ttype, C, S):
    """hybrid [-c] / hybridh / hybridr forward-backward identities: C=1, then
    the `-C2` columns with stride 2 and (tests.py -S) stride 3."""
    if ttype == HERMITIAN and S != C:
        pytest.skip("fftPadHermitian has no stride argument")
    n = 0
    for (L, M, m, D) in cases_1d(ttype, C, S, mult=False):
        check_pad(KIND[ttype], L, M, m, C, S, D)
        n += 1
    assert n >= 8


def _outer(ttype):
    return {COMPLEX: COMPLEX, HERMITIAN: CENTERED, REAL: REAL}[ttype]


def _inner(ttype):
    return {COMPLEX: COMPLEX, HERMITIAN: HERMITIAN, REAL: COMPLEX}[ttype]


@pytest.mark.parametrize("ttype", [COMPLEX, HERMITIAN, REAL])
@pytest.mark.parametrize("extra", [0, 1])
def test_hybridconv_2d_sweep(ttype, extra):
    """hybridconv2 / hybridconvh2 / hybridconvr2: x in detail, y on its basic
    case with every D; extra=1 is the -S run (Sx one word longer)."""
    n = 0
    for (Ly, My, my) in lmm(_inner(ttype), False):
        minS = cq(Ly, 2) if ttype == HERMITIAN else Ly
        for Dy in Ds(_inner(ttype), Ly, My, my, 1):
            if not valid(_inner(ttype), Ly, My, my, 1, 1, Dy):
                continue
            for (Lx, Mx, mx) in lmm(_outer(ttype), True):
                Sx = minS + extra
                for Dx in Ds(_outer(ttype), Lx, Mx, mx, minS):
                    if not valid(_outer(ttype), Lx, Mx, mx, minS, Sx, Dx):
                        continue
                    _check_strided(FAMILY[ttype], [Lx, Ly], [Mx, My], [mx, my], [Dx, Dy],
                                   Sx=Sx)
                    n += 1
    assert n >= 10


@pytest.mark.parametrize("ttype", [COMPLEX, HERMITIAN, REAL])
def test_hybridconv_3d_sweep(ttype):
    """hybridconv3 / hybridconvh3 / hybridconvr3: x in detail, y and z on their
    basic cases (tests.py default)."""
    n = 0
    for (Lz, Mz, mz) in lmm(_inner(ttype), False):
        for Dz in Ds(_inner(ttype), Lz, Mz, mz, 1):
            if not valid(_inner(ttype), Lz, Mz, mz, 1, 1, Dz):
                continue
            Sy = cq(Lz, 2) if ttype == HERMITIAN else Lz
            ymid = CENTERED if ttype == HERMITIAN else COMPLEX
            for (Ly, My, my) in lmm(ymid, False):
                for (Lx, Mx, mx) in lmm(_outer(ttype), True):
                    Cx = Ly * Sy
                    for Dx in Ds(_outer(ttype), Lx, Mx, mx, Cx):
                        if not valid(_outer(ttype), Lx, Mx, mx, Cx, Cx, Dx):
                            continue
                        check(FAMILY[ttype], [Lx, Ly, Lz], [Mx, My, Mz], m=[mx, my, mz],
                              D=[Dx, 1, Dz], I=[0, 0, 0], seed=Lx + Mx + mx)
                        n += 1
    assert n >= 10


def _check_strided(fam, L, M, m, D, Sx):
    """2-D case with an x stride: embed the data, convolve, compare the interior."""
    Lx, Ly = L
    rng = np.random.default_rng(Lx * 131 + Ly + M[0] + m[0])
    if fam == fp.FAMILY_REAL:
        f, g = rng.uniform(-1, 1, L), rng.uniform(-1, 1, L)
        want = O.conv_real(f, g)
        W = Ly
    else:
        W = (Ly + 1) // 2 if fam == fp.FAMILY_HERMITIAN else Ly
        f = rng.uniform(-1, 1, (Lx, W)) + 1j * rng.uniform(-1, 1, (Lx, W))
        g = rng.uniform(-1, 1, (Lx, W)) + 1j * rng.uniform(-1, 1, (Lx, W))
        if fam == fp.FAMILY_HERMITIAN:
            O.symmetrize(L, f)
            O.symmetrize(L, g)
            want = O.conv_hermitian(L, f, g)
        else:
            want = O.conv_complex(f, g)
    conv = fp.HybridConv(L, M, family=fam, m=m, D=D, I=[0, 0], Sx=Sx)
    a = []
    for src in (f, g):
        buf = np.zeros((Lx, Sx), dtype=src.dtype)
        buf[:, :W] = src
        a.append(buf)
    conv.convolve(a)
    assert O.rel_l2(a[0][:, :W], want) < 1e-12, (L, M, m, D, Sx)
    conv.close()
