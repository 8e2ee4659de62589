"""Distributed FFT classes (cpp/mpifftw++.h; reference mpi/mpifftw++.h:37-585,
checked there by mpi/tests/fft2.cc, fft3.cc, fft2r.cc, fft3r.cc against the
serial transforms) on a one-rank communicator: the full code path -- 1-D
passes, pack, NCCL exchange, strided x pass -- against numpy.  The multi-rank
layouts are checked by tests/dist_check.py."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12  # relative to the largest output magnitude (reference tests: 1e-12..1e-10)


@pytest.fixture(scope="module")
def comm():
    import torch
    from fftwpp_b200 import lib
    torch.cuda.init()
    buf = ctypes.create_string_buffer(128)
    assert lib.fftwpp_gpu_comm_unique_id(buf) == 0
    c = ctypes.c_void_p()
    assert lib.fftwpp_gpu_comm_create(0, 1, buf.raw, ctypes.byref(c)) == 0
    yield c
    lib.fftwpp_gpu_comm_destroy(c)


def _rand(shape, seed, cplx):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, shape)
    return a + 1j * rng.uniform(-1, 1, shape) if cplx else a


def _err(got, want):
    return np.max(np.abs(got - want)) / max(1.0, np.max(np.abs(want)))


COMPLEX_SHAPES = [(8, 8), (16, 12), (5, 7), (64, 64), (128, 96), (512, 512), (9, 1), (1, 9),
                  (4, 6, 10), (16, 16, 16), (7, 5, 3), (64, 32, 48), (128, 128, 128), (3, 4, 1)]


@pytest.mark.parametrize("N", COMPLEX_SHAPES, ids=str)
@pytest.mark.parametrize("sign", [-1, 1])
def test_complex_forward_backward(comm, N, sign):
    import torch
    from fftwpp_b200 import dist_conv
    fft = dist_conv.DistributedFFT(N, 0, 1, sign=sign, comm=comm)
    try:
        a = _rand(N, 3, True)
        f = torch.from_numpy(a.copy()).cuda()
        want = np.fft.fftn(a) if sign < 0 else np.fft.ifftn(a) * a.size
        fft.forward(f)
        torch.cuda.synchronize()
        assert _err(f.cpu().numpy(), want) < TOL
        fft.backward(f)
        fft.normalize(f)
        torch.cuda.synchronize()
        assert _err(f.cpu().numpy(), a) < TOL
        # out of place: the input survives a 2-D forward only as scratch, the
        # result is in dst
        g = torch.from_numpy(a.copy()).cuda()
        dst = fft.buffer()
        fft.forward(g, dst)
        torch.cuda.synchronize()
        assert _err(dst.cpu().numpy()[:a.size].reshape(N), want) < TOL
    finally:
        fft.close()


REAL_SHAPES = [(8, 8), (16, 12), (5, 7), (6, 9), (64, 64), (512, 512), (4, 6, 10), (16, 16, 16),
               (7, 5, 3), (5, 3, 8), (64, 32, 48), (128, 128, 128)]


@pytest.mark.parametrize("N", REAL_SHAPES, ids=str)
def test_real_forward_backward(comm, N):
    import torch
    from fftwpp_b200 import dist_conv
    fft = dist_conv.DistributedFFT(N, 0, 1, real=True, comm=comm)
    try:
        a = _rand(N, 5, False)
        f = torch.from_numpy(a.copy()).cuda()
        F = fft.buffer()
        fft.forward(f, F)
        torch.cuda.synchronize()
        want = np.fft.rfftn(a)
        got = F.cpu().numpy()[:want.size].reshape(want.shape)
        assert fft.output_shape() == want.shape
        assert _err(got, want) < TOL
        back = torch.zeros_like(f)
        fft.backward(F, back)
        fft.normalize(back)
        torch.cuda.synchronize()
        assert _err(back.cpu().numpy(), a) < TOL
    finally:
        fft.close()


@pytest.mark.parametrize("N", [(8, 8), (16, 12), (6, 9), (4, 6, 10), (16, 16, 16), (8, 4, 7)], ids=str)
def test_shifted_real_transform_and_denyquist(comm, N):
    """Forward0 = Shift + Forward (Fourier origin at the centre of x, and y in
    3-D) and deNyquist (reference mpifftw++.h:345-372,518-560, mpifftw++.cc:82-101,
    165-186) against the same operations in numpy"""
    import torch
    from fftwpp_b200 import dist_conv
    fft = dist_conv.DistributedFFT(N, 0, 1, real=True, comm=comm)
    try:
        a = _rand(N, 11, False)
        sign = (-1.0) ** np.arange(N[0])
        sign = sign[:, None] if len(N) == 2 else sign[:, None, None] * ((-1.0) ** np.arange(N[1]))[None, :, None]
        f = torch.from_numpy(a.copy()).cuda()
        fft.shift(f)
        torch.cuda.synchronize()
        assert np.array_equal(f.cpu().numpy(), a * sign)          # exact sign flips
        F = fft.buffer()
        fft.forward(f, F)
        fft.denyquist(F)
        torch.cuda.synchronize()
        want = np.fft.rfftn(a * sign)
        want[0] = 0                                               # N[0] is even here
        if len(N) == 2:
            if N[1] % 2 == 0:
                want[:, -1] = 0
        else:
            if N[1] % 2 == 0:
                want[:, 0, :] = 0
            if N[2] % 2 == 0:
                want[:, :, -1] = 0
        got = F.cpu().numpy()[:want.size].reshape(want.shape)
        assert _err(got, want) < TOL
        assert np.all(got[0] == 0)
    finally:
        fft.close()
