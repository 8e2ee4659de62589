"""numpy model of the long-row FFT engine (csrc/tmem_kernels.cu, LongRow16<LG>):
m = 16 x SUB, radix-16 butterfly over a thread's 16 points, one CTA-wide
exchange, then RegFFT<LG-4> sub-transforms inside single warps.  The model
follows the kernel's index arithmetic statement by statement (thread / virtual
thread mapping, padded shared-memory addresses, twiddle table slices), so it
documents and pins the decomposition on CPU:

  * forward() is a DFT (sign +1) in SOME fixed order of the outputs,
  * adjoint() is its exact adjoint (adjoint o forward = m * identity),
  * pointwise products of two forward() results, passed through adjoint(),
    give the circular convolution -- which is all the fused kernel needs.

The GPU kernel itself is checked by tests/test_gpu_long_rows.py."""
import numpy as np
import pytest


def _bfly(a, sign, r):
    u = np.arange(r)
    return a @ np.exp(sign * 2j * np.pi * np.outer(u, u) / r)


def _pos(tau, t, ls):  # RegFFT::pos
    return ((tau >> ls) << (ls + 3)) + (tau & ((1 << ls) - 1)) + (t << ls)


def _pad(p):
    return p + (p >> 3)


class Engine:
    def __init__(self, LG):
        self.N = 1 << LG
        self.SUBLG = LG - 4
        self.SUB = 1 << self.SUBLG
        self.NT = self.SUB
        self.NW = self.NT // 32
        self.TS = self.SUB // 8
        self.NR8, self.REM = self.SUBLG // 3, self.SUBLG % 3
        self.SUBBUF = self.SUB + self.SUB // 8
        self.tid = np.arange(self.NT)
        self.warp, self.lane = self.tid >> 5, self.tid & 31
        tw = np.exp(2j * np.pi * np.arange(self.N // 8) / self.N)   # u=1 block of pass 0 of tw8
        self.wM = tw[:self.SUB]
        self.tabs = []
        for i in range(self.NR8):
            ls = self.SUBLG - 3 * (i + 1)
            self.tabs.append(tw[(16 * np.arange(1 << ls)) << (3 * i)] if ls > 0 else None)

    def ksub(self, a):
        if self.TS == 64:
            return self.warp
        if self.TS == 32:
            return self.warp + self.NW * a
        return (2 * self.warp + a) * (32 // self.TS) + self.lane // self.TS

    def tau(self, a):
        if self.TS == 64:
            return self.lane + 32 * a
        return self.lane if self.TS == 32 else self.lane % self.TS

    def _twid(self, xa, w1, conj):
        w = np.conj(w1) if conj else w1
        return xa * (w[:, None] ** np.arange(8)[None, :])

    def _wex(self, x, buf, ls_from, ls_to):
        for a in range(2):
            for t in range(8):
                buf[self.ksub(a) * self.SUBBUF + _pad(_pos(self.tau(a), t, ls_from))] = x[:, a, t]
        for a in range(2):
            for t in range(8):
                x[:, a, t] = buf[self.ksub(a) * self.SUBBUF + _pad(_pos(self.tau(a), t, ls_to))]

    def _rem(self, x, sign):
        for a in range(2):
            if self.REM == 2:
                x[:, a, 0:4] = _bfly(x[:, a, 0:4], sign, 4)
                x[:, a, 4:8] = _bfly(x[:, a, 4:8], sign, 4)
            elif self.REM == 1:
                for p in range(0, 8, 2):
                    s, d = x[:, a, p] + x[:, a, p + 1], x[:, a, p] - x[:, a, p + 1]
                    x[:, a, p], x[:, a, p + 1] = s, d

    def forward(self, W):
        NT, tid = self.NT, self.tid
        x = np.zeros((NT, 2, 8), complex)
        for a in range(2):
            for t in range(8):
                x[:, a, t] = W[tid + NT * (a + 2 * t)]
        x[:, 0], x[:, 1] = _bfly(x[:, 0], 1, 8), _bfly(x[:, 1], 1, 8)
        x[:, 1] *= np.exp(2j * np.pi * np.arange(8) / 16)[None, :]
        x[:, 0], x[:, 1] = x[:, 0] + x[:, 1], x[:, 0] - x[:, 1]
        for a in range(2):
            for k in range(8):
                x[:, a, k] *= self.wM[tid] ** (k + 8 * a)
        buf = np.zeros(16 * self.SUBBUF, complex)
        for a in range(2):
            for k in range(8):
                buf[(k + 8 * a) * self.SUBBUF + _pad(tid)] = x[:, a, k]
        for a in range(2):
            for t in range(8):
                x[:, a, t] = buf[self.ksub(a) * self.SUBBUF + _pad(self.tau(a) + self.TS * t)]
        for i in range(self.NR8):
            ls = self.SUBLG - 3 * (i + 1)
            for a in range(2):
                x[:, a] = _bfly(x[:, a], 1, 8)
                if ls > 0:
                    x[:, a] = self._twid(x[:, a], self.tabs[i][self.tau(a) & ((1 << ls) - 1)], False)
            ls_next = self.SUBLG - 3 * (i + 2) if i + 1 < self.NR8 else 0
            if i + 1 < self.NR8 or self.REM > 0:
                self._wex(x, buf, ls, ls_next)
        self._rem(x, 1)
        return x

    def adjoint(self, x):
        NT, tid = self.NT, self.tid
        x = x.copy()
        buf = np.zeros(16 * self.SUBBUF, complex)
        self._rem(x, -1)
        for i in range(self.NR8 - 1, -1, -1):
            ls = self.SUBLG - 3 * (i + 1)
            ls_prev = self.SUBLG - 3 * (i + 2) if i + 1 < self.NR8 else 0
            if i + 1 < self.NR8 or self.REM > 0:
                self._wex(x, buf, ls_prev, ls)
            for a in range(2):
                if ls > 0:
                    x[:, a] = self._twid(x[:, a], self.tabs[i][self.tau(a) & ((1 << ls) - 1)], True)
                x[:, a] = _bfly(x[:, a], -1, 8)
        for a in range(2):
            for t in range(8):
                buf[self.ksub(a) * self.SUBBUF + _pad(self.tau(a) + self.TS * t)] = x[:, a, t]
        for a in range(2):
            for k in range(8):
                x[:, a, k] = buf[(k + 8 * a) * self.SUBBUF + _pad(tid)]
        for a in range(2):
            for k in range(8):
                x[:, a, k] *= np.conj(self.wM[tid]) ** (k + 8 * a)
        x[:, 0], x[:, 1] = x[:, 0] + x[:, 1], x[:, 0] - x[:, 1]
        x[:, 1] *= np.conj(np.exp(2j * np.pi * np.arange(8) / 16))[None, :]
        x[:, 0], x[:, 1] = _bfly(x[:, 0], -1, 8), _bfly(x[:, 1], -1, 8)
        out = np.zeros(self.N, complex)
        for a in range(2):
            for t in range(8):
                out[tid + NT * (a + 2 * t)] = x[:, a, t]
        return out


@pytest.mark.parametrize("LG", [11, 12, 13])
def test_engine_model(LG):
    e = Engine(LG)
    rng = np.random.default_rng(LG)
    N = e.N
    W = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    G = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    F = e.forward(W)
    ref = np.fft.ifft(W) * N                                  # the sign +1 DFT
    a = np.sort_complex(np.round(F.ravel(), 6))
    b = np.sort_complex(np.round(ref, 6))
    assert np.max(np.abs(a - b)) < 1e-5                       # same values, some order
    assert np.max(np.abs(e.adjoint(F) / N - W)) < 1e-12
    h = e.adjoint(e.forward(W) * e.forward(G)) / N
    want = np.fft.ifft(np.fft.fft(W) * np.fft.fft(G))
    assert np.max(np.abs(h - want)) / np.max(np.abs(want)) < 1e-13
    # every shared-memory slot of the CTA-wide exchange is written exactly once
    slots = np.concatenate([(k + 8 * a2) * e.SUBBUF + _pad(e.tid) for a2 in range(2) for k in range(8)])
    assert len(np.unique(slots)) == N
    reads = np.concatenate([e.ksub(a2) * e.SUBBUF + _pad(e.tau(a2) + e.TS * t)
                            for a2 in range(2) for t in range(8)])
    assert np.array_equal(np.sort(reads), np.sort(slots))     # and read exactly once
