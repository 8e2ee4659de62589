"""Parity at the FULL sizes of the BASELINE.json configs (cfg1 is in
test_gpu_conv.py::test_cfg1_closed_form_full_size), through the C ABI, against
the oracle on seeded inputs; tolerance 1e-12*log2(N) (north_star)."""
import numpy as np
import pytest

import fftwpp_b200 as fp
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def test_cfg2_conv2d_4096_full_size():
    """configs[1]: hybridconv2, 2-D complex 4096 x 4096 (tests/hybridconv2.cc)."""
    n = 4096
    rng = np.random.default_rng(1234)
    f, g = crand(rng, n, n), crand(rng, n, n)
    want = O.conv_complex(f, g)
    conv = fp.HybridConv([n, n], [2 * n, 2 * n])
    a = [f, g]
    conv.convolve(a)
    padded = [conv.params(d)["m"] * conv.params(d)["q"] for d in range(2)]
    assert O.rel_l2(a[0], want) < O.tolerance(*padded)
    conv.close()


def test_cfg3_hermitian_256_full_size():
    """configs[2]: hybridconvh3, centred Hermitian 256^3, M=3L/2
    (tests/hybridconvh3.cc; inputs symmetrised with the PRODUCT's
    HermitianSymmetrizeXY as the reference driver does, :72)."""
    import ctypes
    n = 256
    L = [n, n, n]
    shp = (n, n, n // 2)
    rng = np.random.default_rng(1235)
    f, g = crand(rng, *shp), crand(rng, *shp)
    for a in (f, g):
        fp.lib.fftwpp_HermitianSymmetrizeXY(n // 2, n // 2, n // 2, n // 2, n // 2,
                                            ctypes.c_void_p(a.ctypes.data))
    fo, go = f.copy(), g.copy()
    assert np.array_equal(O.symmetrize(L, fo.copy()), fo)  # already symmetric
    want = O.conv_hermitian(L, fo, go)
    conv = fp.HybridConv(L, [384, 384, 384], family=fp.FAMILY_HERMITIAN)
    a = [f, g]
    conv.convolve(a)
    padded = [conv.params(d)["m"] * conv.params(d)["q"] for d in range(3)]
    assert O.rel_l2(a[0], want) < O.tolerance(*padded)
    conv.close()


def test_cfg4_real_512_full_size():
    """configs[3] on one GPU: hybridconvr3 512^3 (tests/hybridconvr3.cc); the
    whole output volume is compared on z-pencils (corners, edges, interior)
    and through a size-independent property (linearity in the first input)."""
    import torch
    n = 512
    rng = np.random.default_rng(1236)
    f, g = rng.uniform(-1, 1, (n, n, n)), rng.uniform(-1, 1, (n, n, n))
    pts = [(i, j) for i in (0, 1, 200, n - 1) for j in (0, 3, 257, n - 1)]
    want = O.conv_real_pencils(f, g, pts)
    conv = fp.HybridConv([n] * 3, [2 * n] * 3, family=fp.FAMILY_REAL)
    d = [torch.from_numpy(f).cuda(), torch.from_numpy(g).cuda()]
    conv.convolve(d)
    torch.cuda.synchronize()
    h1 = d[0]
    num = den = 0.0
    for (i, j) in pts:
        got = h1[i, j].cpu().numpy()
        num += np.sum((got - want[(i, j)]) ** 2)
        den += np.sum(want[(i, j)] ** 2)
    assert np.sqrt(num / den) < O.tolerance(2 * n, 2 * n, 2 * n)
    # linearity: conv(2 f - f2, g) = 2 conv(f,g) - conv(f2,g), whole volume on the device
    f2 = rng.uniform(-1, 1, (n, n, n))
    d2 = [torch.from_numpy(f2).cuda(), torch.from_numpy(g).cuda()]
    conv.convolve(d2)
    d3 = [torch.from_numpy(2.0 * f - f2).cuda(), torch.from_numpy(g).cuda()]
    conv.convolve(d3)
    torch.cuda.synchronize()
    comb = 2.0 * h1 - d2[0]
    err = (torch.linalg.vector_norm(d3[0] - comb) / torch.linalg.vector_norm(comb)).item()
    assert err < 1e-13
    conv.close()


def test_cfg5_batched_rows_full_size():
    """configs[4]: 4096 independent 1-D complex convolutions L=8192, M=16384,
    contiguous rows, one batched launch; every row checked."""
    import torch
    import scipy.fft as sf
    L, rows = 8192, 4096
    rng = np.random.default_rng(1237)
    f, g = crand(rng, rows, L), crand(rng, rows, L)
    F = sf.fft(f, 2 * L, axis=1, workers=-1)
    F *= sf.fft(g, 2 * L, axis=1, workers=-1)
    want = sf.ifft(F, axis=1, workers=-1)[:, :L]
    conv = fp.HybridConv([L], [2 * L])
    d = [torch.from_numpy(f).cuda(), torch.from_numpy(g).cuda()]
    conv.convolve_rows(d, rows, L)
    torch.cuda.synchronize()
    got = d[0].cpu().numpy()
    per_row = np.sqrt(np.sum(np.abs(got - want) ** 2, axis=1) / np.sum(np.abs(want) ** 2, axis=1))
    assert per_row.max() < O.tolerance(2 * L)
    # ragged batch: 3 rows of the same plan
    d = [torch.from_numpy(f[:3].copy()).cuda(), torch.from_numpy(g[:3].copy()).cuda()]
    conv.convolve_rows(d, 3, L)
    torch.cuda.synchronize()
    assert O.rel_l2(d[0].cpu().numpy(), want[:3]) < O.tolerance(2 * L)
    conv.close()


def test_cfg5_interleaved_layout_full_size():
    """configs[4], layout (ii): fftPad(8192,16384,C=4096,S=4096) "Many"
    forward/backward identity on all residues (tests/hybrid.cc -C -S)."""
    import torch
    L, C = 8192, 4096
    rng = np.random.default_rng(1238)
    P = fp.Pad(fp.KIND_COMPLEX, L, 2 * L, C, C, 0, 0, -1, A=1, B=1)
    f = crand(rng, L, 64)
    full = np.zeros((L, C), dtype=np.complex128)
    full[:, :64] = f
    full[:, -1] = f[:, 0]
    dev = torch.from_numpy(full).cuda()
    out = torch.zeros((L, C), dtype=torch.complex128, device="cuda")
    F = torch.zeros(P.outputSize, dtype=torch.complex128, device="cuda")
    N = P.m * P.q
    F2 = O.padded_dft(O.KIND_COMPLEX, L, N, f[:, :4])
    for r in P.residue_calls():
        P.forward(dev, r, F)
        torch.cuda.synchronize()
        Fh = F.cpu().numpy()
        nout = P.noutputs(r)
        for k in list(range(0, nout, max(1, nout // 97))) + [nout - 1]:
            idx = P.index(r, k)
            assert np.allclose(Fh[k * C:k * C + 4], F2[idx], rtol=0, atol=1e-10), (r, k)
        P.backward(F, out, r)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert O.rel_l2(got[:, :64], N * f) < O.tolerance(N)
    assert O.rel_l2(got[:, -1], N * f[:, 0]) < O.tolerance(N)
    P.close()
