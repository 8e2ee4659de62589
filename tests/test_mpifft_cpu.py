"""CPU, one process playing every rank: the splits and exchange tables of the
distributed FFT classes (cpp/mpifftw++.h; reference mpi/mpifftw++.h:37-585)
must realise x x Y [x Z] -> X x y [x Z] for every rank count up to 8, including
ranks that own no x rows and / or no y rows.  The local 1-D passes are numpy's
here (the GPU passes are checked by tests/test_gpu_mpifft.py and
tests/dist_check.py); what is tested is the data flow: local passes, pack by
destination, all-to-all by the product's tables, strided x pass, and the
inverse -- the steps of fft2dMPI::Forward / Backward (mpifftw++.cc:7-80)."""
import ctypes

import numpy as np
import pytest

from fftwpp_b200 import dist_conv

SHAPES = [(16, 12), (9, 7), (9, 3), (3, 10), (5, 5), (8, 6, 10), (5, 9, 7), (3, 2, 4), (16, 16, 6)]


def _handles(N, world, real):
    return [dist_conv.DistributedFFT(N, r, world, real=real, comm=ctypes.c_void_p())
            for r in range(world)]


def _alltoall(tables, send):
    """send[r]: bytes of rank r's send buffer; returns the receive buffers"""
    world = len(send)
    recv = [np.zeros(max(1, sum(tables[r][2])), dtype=np.uint8) for r in range(world)]
    for r in range(world):
        sc, sd, _, _ = tables[r]
        for p in range(world):
            rc, rd = tables[p][2], tables[p][3]
            assert sc[p] == rc[r], "rank %d sends %d bytes to %d, which expects %d" % (r, sc[p], p, rc[r])
            recv[p][rd[r]:rd[r] + rc[r]] = send[r][sd[p]:sd[p] + sc[p]]
    return recv


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("N", SHAPES, ids=str)
@pytest.mark.parametrize("real", [False, True])
def test_data_flow(N, world, real):
    if real and N[-1] < 2:
        pytest.skip("real transforms need a last dimension of at least 2")
    rng = np.random.default_rng(sum(N) + world)
    full = rng.uniform(-1, 1, N) + (0 if real else 1j * rng.uniform(-1, 1, N))
    want = np.fft.rfftn(full) if real else np.fft.fftn(full)
    ffts = _handles(N, world, real)
    try:
        sp = [f.split for f in ffts]
        X, Y, Z = sp[0]["X"], sp[0]["Y"], sp[0]["Z"]
        assert (X, Y, Z) == (want.shape + (1,))[:3]
        # every row / column is owned exactly once, in rank order
        assert sum(s["x"] for s in sp) == X and sum(s["y"] for s in sp) == Y
        for r in range(world):
            assert (sp[r]["x"], sp[r]["x0"]) == dist_conv.local_dimension(X, r, world)
            assert (sp[r]["y"], sp[r]["y0"]) == dist_conv.local_dimension(Y, r, world)
            assert ffts[r].words >= max(X * sp[r]["y"], sp[r]["x"] * Y) * Z
        # ---- forward: local passes on x x Y x Z, pack, exchange, x pass ----
        tables = [f.exchange_table(1) for f in ffts]
        send = []
        for r in range(world):
            s = sp[r]
            loc = full[s["x0"]:s["x0"] + s["x"]]
            if real:
                T = np.fft.rfft(loc, axis=-1)
                if len(N) == 3:
                    T = np.fft.fft(T, axis=1)
            else:
                T = np.fft.fftn(loc, axes=tuple(range(1, len(N))))
            T = T.reshape(s["x"], Y, Z)
            buf = np.zeros(max(1, sum(tables[r][0])), dtype=np.uint8)
            for p in range(world):                      # [x rows][py][Z] per destination
                py, py0 = sp[p]["y"], sp[p]["y0"]
                blk = np.ascontiguousarray(T[:, py0:py0 + py, :]).reshape(-1).view(np.uint8)
                assert blk.size == tables[r][0][p]
                buf[tables[r][1][p]:tables[r][1][p] + blk.size] = blk
            send.append(buf)
        recv = _alltoall(tables, send)
        out = []
        for r in range(world):
            s = sp[r]
            F = recv[r][:X * s["y"] * Z * 16].view(np.complex128).reshape(X, s["y"], Z)
            F = np.fft.fft(F, axis=0)
            w = want.reshape(X, Y, Z)[:, s["y0"]:s["y0"] + s["y"], :]
            assert np.allclose(F, w, atol=1e-10), "forward mismatch on rank %d of %d" % (r, world)
            out.append(F)
        # ---- backward: x pass, exchange, unpack by source, local passes ----
        tables = [f.exchange_table(0) for f in ffts]
        send = [np.ascontiguousarray(np.fft.ifft(out[r], axis=0) * X).reshape(-1).view(np.uint8)
                if out[r].size else np.zeros(1, dtype=np.uint8) for r in range(world)]
        recv = _alltoall(tables, send)
        for r in range(world):
            s = sp[r]
            T = np.zeros((s["x"], Y, Z), dtype=np.complex128)
            for p in range(world):
                py, py0 = sp[p]["y"], sp[p]["y0"]
                n = s["x"] * py * Z * 16
                assert n == tables[r][2][p]
                T[:, py0:py0 + py, :] = recv[r][tables[r][3][p]:tables[r][3][p] + n].view(
                    np.complex128).reshape(s["x"], py, Z)
            loc = full[s["x0"]:s["x0"] + s["x"]]
            wantT = (np.fft.rfft(loc, axis=-1) if real else loc.astype(np.complex128))
            if real and len(N) == 3:
                wantT = np.fft.fft(wantT, axis=1)
            if not real:
                wantT = np.fft.fftn(loc, axes=tuple(range(1, len(N))))
            assert np.allclose(T / X, wantT.reshape(s["x"], Y, Z), atol=1e-10), \
                "backward mismatch on rank %d of %d" % (r, world)
    finally:
        for f in ffts:
            f.close()
