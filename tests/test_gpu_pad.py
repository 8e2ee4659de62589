"""GPU parity of the residue-level padded FFTs (fft->forward / fft->backward)
against the explicit padded DFT, walked through index(r,k) exactly as the
reference's tests/hybrid.cc:72-167, hybridh.cc and hybridr.cc:72-141 do.
Calls go through the C ABI (lib_fftwpp.so) with host buffers."""
import numpy as np
import pytest

import fftwpp_b200 as fp
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-12  # reference tests/tests.py:223-224


def _valid_D(kind, m, p, q, n, S, C):
    out = []
    for D in range(1, n + 1):
        if kind in (0, 1):
            ok = (D == 1) if q == 1 else (D == 1 or (S == 1 and ((D < n and D % 2 == 0) or D == n)))
            if kind == 1:
                ok = ok and (q == 1 or p % 2 == 0)
        elif kind == 2:
            ok = (D == 1 and q == 1) or (D == 2 and p % 2 == 0 and (p == 2 or C == 1))
        else:
            ok = ((n % 2 == 1 or (p % 2 == 0 or p <= 2)) and (q % 2 == 1 or m % 2 == 0)
                  and (D == 1 or (S == 1 and ((D < (n - 1) // 2 and D % 2 == 0)
                                              or D == (n - 1) // 2))))
        if ok:
            out.append(D)
    return out


def cases():
    out = []
    for kind in (0, 1, 2, 3):
        for L in (3, 5, 8, 12):
            for M in sorted(set([2 * L, (3 * L + 1) // 2, 5 * L // 2, 4 * L])):
                for m in sorted(set([M, L + 1, L, (L + 1) // 2, max(2, L // 4),
                                     O.nextfftsize(M)])):
                    for (C, S) in ((1, 1), (3, 3), (2, 3)):
                        if kind == 2 and S != C:
                            continue
                        p, n, q = O.parameters(L, M, m, kind in (1, 2))
                        if q * m < M:
                            continue
                        if kind == 1 and q > 1 and p % 2:
                            continue
                        for D in _valid_D(kind, m, p, q, n, S, C):
                            out.append((kind, L, M, m, C, S, D))
    return out


def _input(kind, L, C, S, rng):
    Lin = (L + 1) // 2 if kind == 2 else L
    if kind == 3:
        f = np.zeros((Lin, S))
        f[:, :C] = rng.uniform(-1, 1, (Lin, C))
    else:
        f = np.zeros((Lin, S), dtype=np.complex128)
        f[:, :C] = rng.uniform(-1, 1, (Lin, C)) + 1j * rng.uniform(-1, 1, (Lin, C))
        if kind == 2:
            f[0] = f[0].real
    return f


@pytest.mark.parametrize("kind,L,M,m,C,S,D", cases())
def test_forward_backward(kind, L, M, m, C, S, D):
    check_pad(kind, L, M, m, C, S, D)


def check_pad(kind, L, M, m, C, S, D):
    """Forward Error / Backward Error of tests/hybrid.cc, hybridh.cc, hybridr.cc."""
    rng = np.random.default_rng(1234 + 7 * L + M + 13 * m)
    pad = fp.Pad(kind, L, M, C, S, m, D, 0, A=1, B=1)
    N = pad.paddedSize
    f = _input(kind, L, C, S, rng)
    F2 = O.padded_dft(kind, L, N, f[:, :C])
    err = norm = 0.0
    h = np.zeros_like(f)
    for r in pad.residue_calls():
        F = pad.forward(f, r)
        G = np.zeros_like(F)
        if kind == 2:
            Fr = F.view(np.float64)
            Gr = G.view(np.float64)
            stride = pad.noutputs(0)
            for d in range(pad.D0 if r == 0 else pad.D):
                base = 2 * pad.b * d
                for k in range(stride):
                    i = pad.index(r, k + stride * d)
                    val = F2[i]
                    got = Fr[base + C * k: base + C * k + C]
                    err += np.sum(np.abs(got - val) ** 2)
                    norm += np.sum(np.abs(val) ** 2)
                    Gr[base + C * k: base + C * k + C] = val
        else:
            for k in range(pad.noutputs(r)):
                i = pad.index(r, k)
                val = (np.array([O.real_spectrum_at(F2[:, c], N, i) for c in range(C)])
                       if kind == 3 else F2[i])
                got = F[S * k: S * k + C]
                err += np.sum(np.abs(got - val) ** 2)
                norm += np.sum(np.abs(val) ** 2)
                G[S * k: S * k + C] = val
        # backward from the exact spectrum must rebuild N*f
        pad.backward(G, h, r)
    assert np.sqrt(err / norm) < TOL
    scale = 1.0 / pad.normalization
    e2 = O.rel_l2(h[:, :C] * scale, f[:, :C])
    assert e2 < TOL
    pad.close()
