"""CPU: pin the oracle (oracle/oracle.py, oracle/direct_oracle.c) against the
golden fixtures produced by the reference itself (tests/golden/make_golden.py)
and, when oracle/_ref is built, against the live reference."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")

with open(os.path.join(G, "cases_meta.json")) as fh:
    META = json.load(fh)
CONV = np.load(os.path.join(G, "conv_cases.npz"))
FWD = np.load(os.path.join(G, "forward_cases.npz"))


def _oracle_conv(case, f, g):
    fam = case["family"]
    if fam == 0:
        return O.conv_complex(f, g)
    if fam == 2:
        return O.conv_real(f, g)
    return O.conv_hermitian(case["L"], f, g)


@pytest.mark.parametrize("case", META["conv"], ids=[c["name"] for c in META["conv"]])
def test_numpy_oracle_matches_reference_hybrid_and_direct(case):
    n = case["name"]
    f, g = CONV[n + "_f"], CONV[n + "_g"]
    want = _oracle_conv(case, f, g)
    # reference hybrid (FFTW++ logic on the shim FFT) and reference direct sums
    assert O.rel_l2(CONV[n + "_hybrid"], want) < 1e-13
    assert O.rel_l2(CONV[n + "_direct"], want) < 1e-13


@pytest.mark.parametrize("case", META["conv"], ids=[c["name"] for c in META["conv"]])
def test_c_direct_oracle_matches_reference_direct(case):
    n = case["name"]
    f, g = CONV[n + "_f"], CONV[n + "_g"]
    kind = {0: "complex", 1: "hermitian", 2: "real"}[case["family"]]
    got = O.direct(kind, f, g, case["L"])
    assert O.rel_l2(got, CONV[n + "_direct"]) < 1e-14


@pytest.mark.parametrize("L", [8, 7])
def test_centered_direct(L):
    f, g = CONV["cen%d_f" % L], CONV["cen%d_g" % L]
    want = CONV["cen%d_direct" % L]
    assert O.rel_l2(O.conv_centered1(f, g), want) < 1e-13
    assert O.rel_l2(O.direct("centered", f, g), want) < 1e-14


def test_closed_form():
    # tests/hybridconv.cc:50-55,65-69 (`-a`): usable at any L
    for L in (8, 1000, 1 << 16):
        f, g, h = O.closed_form_1d(L)
        assert O.rel_l2(O.conv_complex(f, g), h) < 1e-12


@pytest.mark.parametrize("case", META["forward"], ids=[c["name"] for c in META["forward"]])
def test_padded_dft_and_index_match_reference_forward(case):
    """Residue-level oracle: explicit padded DFT walked through the restated
    index() must reproduce the reference's fft->forward output."""
    kind, L, M, C, S, m, D, I = case["args"]
    centered = kind in (1, 2)
    p, n, q = O.parameters(L, M, m, centered)
    N = m * q
    f = FWD[case["name"] + "_f"]
    F2 = O.padded_dft(kind, L, N, f[:, :C])
    if kind == 3:
        nres = n
        D0 = ((n - 1) // 2) % D or D
    else:
        D0 = n % D or D
    e = m // 2 + 1
    for r in case["calls"]:
        F = FWD["%s_F%d" % (case["name"], r)]
        if kind == 2:
            b = F.shape[0] // D
            Fr = F.view(np.float64)
            blocks = D0 if r == 0 else D
            stride = m * (1 if q == 1 else p // 2)  # blocksize (convolve.h:786-793)
            for d in range(blocks):
                for k in range(stride):
                    i = O.index_complex(r, k + stride * d, m=m, p=p, q=q, n=n, D=D, D0=D0,
                                        centered=True)
                    got = Fr[2 * b * d + C * k: 2 * b * d + C * k + C]
                    assert np.allclose(got, F2[i], rtol=0, atol=1e-12 * max(1, np.abs(F2).max()))
        elif kind == 3:
            nout = O.real_blocksize(r, m=m, p=p, n=n)
            if r > 0 and 2 * r != n:
                nout *= D0 if r == 1 else D
            for k in range(nout):
                i = O.index_real(r, k, m=m, p=p, q=q, n=n)
                val = np.array([O.real_spectrum_at(F2[:, c], N, i) for c in range(C)])
                assert np.allclose(F[S * k: S * k + C], val, rtol=0,
                                   atol=1e-12 * max(1, np.abs(F2).max()))
        else:
            P = 1 if p == 2 else (p // 2 if centered else p)
            if q == 1:
                P = 1
            nout = (m if q == 1 else m * P) * (1 if q == 1 else (D0 if r == 0 else D))
            for k in range(nout):
                i = O.index_complex(r, k, m=m, p=p, q=q, n=n, D=D, D0=D0, centered=centered)
                assert np.allclose(F[S * k: S * k + C], F2[i], rtol=0,
                                   atol=1e-12 * max(1, np.abs(F2).max()))


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_live_reference_agrees_with_oracle():
    rng = np.random.default_rng(3)
    f, g = rng.uniform(-1, 1, (6, 5, 4)), rng.uniform(-1, 1, (6, 5, 4))
    c = R.RefConv([6, 5, 4], [12, 10, 8], family=2)
    a = [f.copy(), g.copy()]
    c.convolve(a)
    c.close()
    assert O.rel_l2(a[0], O.conv_real(f, g)) < 1e-13
    assert O.rel_l2(R.direct_real(f, g), O.direct("real", f, g)) < 1e-14
