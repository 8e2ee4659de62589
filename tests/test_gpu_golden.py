"""GPU: the CUDA path against the committed golden fixtures (outputs of the
reference itself), through the C ABI with host buffers."""
import json
import os

import numpy as np
import pytest

import fftwpp_b200 as fp
from oracle import oracle as O

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
with open(os.path.join(G, "cases_meta.json")) as fh:
    META = json.load(fh)
CONV = np.load(os.path.join(G, "conv_cases.npz"))
FWD = np.load(os.path.join(G, "forward_cases.npz"))


@pytest.mark.parametrize("case", META["conv"], ids=[c["name"] for c in META["conv"]])
def test_convolution_matches_reference_output(case):
    n = case["name"]
    f, g = CONV[n + "_f"], CONV[n + "_g"]
    # same (m,D,I) the reference used, so even the residue schedule matches
    m = [p["m"] for p in case["params"]]
    D = [p["D"] for p in case["params"]]
    ok = True
    try_forced = all(not (p["q"] > 1 and p["p"] > 2 and case["family"] != 0)
                     for p in case["params"])
    kw = dict(m=m, D=D, I=[0] * len(m)) if try_forced else {}
    conv = fp.HybridConv(case["L"], case["M"], family=case["family"], **kw)
    a = [np.ascontiguousarray(f.copy()), np.ascontiguousarray(g.copy())]
    conv.convolve(a)
    assert O.rel_l2(a[0], CONV[n + "_hybrid"]) < 1e-12
    assert O.rel_l2(a[0], CONV[n + "_direct"]) < 1e-12
    assert ok


@pytest.mark.parametrize("case", META["forward"], ids=[c["name"] for c in META["forward"]])
def test_forward_matches_reference_layout(case):
    kind, L, M, C, S, m, D, I = case["args"]
    pad = fp.Pad(kind, L, M, C, S, m, D, I, A=1, B=1)
    f = FWD[case["name"] + "_f"]
    for r in case["calls"]:
        want = FWD["%s_F%d" % (case["name"], r)]
        got = pad.forward(np.ascontiguousarray(f.copy()), r)
        # compare only the positions the reference defines (column c < C of
        # each produced row); gaps of strided layouts are unspecified
        if kind == 2:
            stride = pad.noutputs(0)
            b = pad.b
            for d in range(pad.D0 if r == 0 else pad.D):
                gr = got.view(np.float64)[2 * b * d: 2 * b * d + C * stride]
                wr = want.view(np.float64)[2 * b * d: 2 * b * d + C * stride]
                assert np.allclose(gr, wr, rtol=0, atol=1e-12 * max(1, np.abs(wr).max()))
        else:
            for k in range(pad.noutputs(r)):
                gk = got[S * k: S * k + C]
                wk = want[S * k: S * k + C]
                assert np.allclose(gk, wk, rtol=0, atol=1e-12 * max(1, np.abs(want).max()))
    pad.close()
