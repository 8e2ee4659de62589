"""Rows of m = 8192 and m = 4096 points on the tensor-memory kernel (fast_conv_rows_long,
csrc/tmem_kernels.cu): the fully fused residue loop of Convolution::convolveRaw
(reference convolve.cc:7513-7575) for 4096 < L <= 8192, against the numpy oracle.
Covers full rows, zero-padded rows (L < m), odd L, the correlation multiplier,
ragged batches, row strides, 2-D use as the y pass, and the selected
parameters."""
import numpy as np
import pytest

import fftwpp_b200 as fp
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def rows_want(f, g, M, corr=False):
    F = np.fft.ifft(f, M, axis=1) * M
    G = np.fft.ifft(g, M, axis=1) * M
    h = np.fft.fft(F * (np.conj(G) if corr else G), axis=1) / M
    return h[:, :f.shape[1]]


@pytest.mark.parametrize("L,M,m,force", [(8192, 16384, 8192, False), (8192, 12289, 8192, True),
                                         (5000, 10000, 8192, True), (4097, 8194, 8192, True),
                                         (8191, 16382, 8192, True), (6144, 16384, 8192, True),
                                         (4096, 8192, 4096, False), (3000, 6000, 4096, True),
                                         (2049, 4098, 4096, True), (4095, 8190, 4096, True),
                                         (2048, 4096, 2048, False), (1500, 3000, 2048, True),
                                         (1025, 2050, 2048, True)])
@pytest.mark.parametrize("mult", [fp.MULT_BINARY, fp.MULT_CORRELATION])
def test_long_rows(L, M, m, force, mult):
    import torch
    rng = np.random.default_rng(L + mult)
    rows = 5
    conv = fp.HybridConv([L], [M], m=[m] if force else None, mult=mult)
    p = conv.params(0)
    assert (p["m"], p["p"], p["q"]) == (m, 1, 2)
    f, g = crand(rng, rows, L), crand(rng, rows, L)
    want = rows_want(f, g, 2 * m, corr=mult == fp.MULT_CORRELATION)
    d = [torch.from_numpy(f.copy()).cuda(), torch.from_numpy(g.copy()).cuda()]
    before = fp.lib.fftwpp_gpu_launch_count()
    conv.convolve_rows(d, rows, L)
    torch.cuda.synchronize()
    assert fp.lib.fftwpp_gpu_launch_count() - before == 1     # one fused launch
    got = d[0].cpu().numpy()
    per_row = np.sqrt(np.sum(np.abs(got - want) ** 2, axis=1) / np.sum(np.abs(want) ** 2, axis=1))
    assert per_row.max() < O.tolerance(2 * m)
    assert np.array_equal(d[1].cpu().numpy(), g)              # second input untouched
    conv.close()


def test_long_rows_many_rows_and_stride():
    """more rows than CTAs (persistent loop) and a row stride > L"""
    import torch
    rng = np.random.default_rng(77)
    L, rows, rs = 8192, 333, 8192 + 64
    conv = fp.HybridConv([L], [2 * L])
    f, g = crand(rng, rows, rs), crand(rng, rows, rs)
    want = rows_want(f[:, :L], g[:, :L], 2 * L)
    d = [torch.from_numpy(f.copy()).cuda(), torch.from_numpy(g.copy()).cuda()]
    conv.convolve_rows(d, rows, rs)
    torch.cuda.synchronize()
    got = d[0].cpu().numpy()
    assert O.rel_l2(got[:, :L], want) < O.tolerance(2 * L)
    assert np.array_equal(got[:, L:], f[:, L:])               # the gap is not written
    conv.close()


def test_long_rows_single_convolution_closed_form():
    """1-D convolve() entry (one row) against the reference's closed form
    (tests/hybridconv.cc: f_j = g_j = j+1 variants are covered elsewhere;
    here the oracle convolution)."""
    import torch
    rng = np.random.default_rng(78)
    L = 8192
    conv = fp.HybridConv([L], [2 * L])
    f, g = crand(rng, L), crand(rng, L)
    want = O.conv_complex(f, g)
    d = [torch.from_numpy(f.copy()).cuda(), torch.from_numpy(g.copy()).cuda()]
    conv.convolve(d)
    torch.cuda.synchronize()
    assert O.rel_l2(d[0].cpu().numpy(), want) < O.tolerance(2 * L)
    conv.close()


def test_long_rows_as_y_pass_of_2d():
    """2-D complex convolution whose contiguous dimension has L = 8192"""
    import torch
    rng = np.random.default_rng(79)
    Lx, Ly = 12, 8192
    conv = fp.HybridConv([Lx, Ly], [2 * Lx, 2 * Ly])
    assert conv.params(1)["m"] == 8192
    f, g = crand(rng, Lx, Ly), crand(rng, Lx, Ly)
    want = O.conv_complex(f, g)
    d = [torch.from_numpy(f.copy()).cuda(), torch.from_numpy(g.copy()).cuda()]
    conv.convolve(d)
    torch.cuda.synchronize()
    assert O.rel_l2(d[0].cpu().numpy(), want) < O.tolerance(2 * Lx, 2 * Ly)
    conv.close()
