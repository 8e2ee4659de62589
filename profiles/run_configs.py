"""Timing + parity spot checks of the five BASELINE configs on one B200
(device-resident data, CUDA events).  Writes one JSON line per config."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftwpp_b200 as fp  # noqa: E402
from oracle import oracle as O  # noqa: E402

PEAK = 6536.4


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def report(name, ms, model_bytes, err, params):
    gbs = model_bytes / 1e9 / (ms / 1e3)
    print(json.dumps({"config": name, "ms": ms, "conv_per_s": 1e3 / ms,
                      "sweep_model_GB": model_bytes / 1e9, "GBps": gbs,
                      "frac_of_measured_hbm": gbs / PEAK, "rel_l2_vs_oracle": err,
                      "params": params}))


def main():
    rng = np.random.default_rng(1234)
    # cfg1: 1-D complex L=2^20, M=2^21
    L = 1 << 20
    c = fp.HybridConv([L], [2 * L])
    f = rng.uniform(-1, 1, L) + 1j * rng.uniform(-1, 1, L)
    g = rng.uniform(-1, 1, L) + 1j * rng.uniform(-1, 1, L)
    a = [f.copy(), g.copy()]
    c.convolve(a)
    err = O.rel_l2(a[0], O.conv_complex(f, g))
    d = [torch.zeros(L, dtype=torch.complex128, device="cuda") for _ in range(2)]
    ms = timeit(lambda: c.convolve(d, normalized=False), 50)
    report("cfg1 1-D complex L=2^20 M=2^21", ms, 15 * L * 16, err, [c.params(0)])
    # cfg1 is the reference's own CPU-runnable case: time it on the host cores
    # beside the GPU number (oracle/_ref, its optimizer's choice, all threads)
    try:
        from oracle import ref as R
        if R.available():
            cores = min(int(R.lib().ref_get_max_threads()), os.cpu_count() or 1)
            rc = R.RefConv([L], [2 * L], family=0, threads=cores)
            t = sorted(rc.time(7)[2:])
            print(json.dumps({"config": "cfg1 on the host CPU (reference, oracle/_ref)",
                              "ms": 1e3 * t[len(t) // 2], "cores": cores,
                              "params": [rc.params(0)]}))
            rc.close()
    except Exception as e:  # the CPU line is informative only
        print(json.dumps({"config": "cfg1 on the host CPU", "error": repr(e)}))

    # cfg2: 2-D complex 4096^2
    n = 4096
    c = fp.HybridConv([n, n], [2 * n, 2 * n])
    sub = 256  # parity on a sub-problem of the same kernels is checked in tests
    d = [torch.zeros((n, n), dtype=torch.complex128, device="cuda") for _ in range(2)]
    ms = timeit(lambda: c.convolve(d, normalized=False), 5)
    f = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
    g = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
    a = [f.copy(), g.copy()]
    c.convolve(a)
    err = O.rel_l2(a[0], O.conv_complex(f, g))
    report("cfg2 2-D complex 4096^2", ms, 15 * n * n * 16, err, [c.params(0), c.params(1)])
    del d, a, f, g

    # cfg3: 3-D centred Hermitian 256^3 (M=384)
    n = 256
    c = fp.HybridConv([n, n, n], [384, 384, 384], family=fp.FAMILY_HERMITIAN)
    shp = (n, n, n // 2)
    f = rng.uniform(-1, 1, shp) + 1j * rng.uniform(-1, 1, shp)
    g = rng.uniform(-1, 1, shp) + 1j * rng.uniform(-1, 1, shp)
    O.symmetrize([n, n, n], f)
    O.symmetrize([n, n, n], g)
    a = [f.copy(), g.copy()]
    c.convolve(a)
    err = O.rel_l2(a[0], O.conv_hermitian([n, n, n], f, g))
    d = [torch.zeros(shp, dtype=torch.complex128, device="cuda") for _ in range(2)]
    ms = timeit(lambda: c.convolve(d, normalized=False), 5)
    report("cfg3 3-D centred Hermitian 256^3", ms, 25.5 * n * n * (n // 2) * 16, err,
           [c.params(i) for i in range(3)])
    del d

    # cfg4 is bench.py itself; parity at 512^3 against the numpy oracle here
    n = 512
    c = fp.HybridConv([n] * 3, [2 * n] * 3, family=fp.FAMILY_REAL)
    f = rng.uniform(-1, 1, (n, n, n))
    g = rng.uniform(-1, 1, (n, n, n))
    t0 = time.time()
    want = O.conv_real(f, g)
    a = [f.copy(), g.copy()]
    c.convolve(a)
    err = O.rel_l2(a[0], want)
    print(json.dumps({"config": "cfg4 3-D real 512^3 parity vs numpy oracle",
                      "rel_l2_vs_oracle": err, "tolerance": O.tolerance(1024, 1024, 1024),
                      "oracle_seconds": time.time() - t0}))
    del a, f, g, want

    # cfg5: 4096 independent 1-D complex L=8192, M=16384
    L, rows = 8192, 4096
    c = fp.HybridConv([L], [2 * L])
    d = [torch.zeros((rows, L), dtype=torch.complex128, device="cuda") for _ in range(2)]
    ms = timeit(lambda: c.convolve_rows(d, rows, L, normalized=False), 10)
    f = rng.uniform(-1, 1, (8, L)) + 1j * rng.uniform(-1, 1, (8, L))
    g = rng.uniform(-1, 1, (8, L)) + 1j * rng.uniform(-1, 1, (8, L))
    td = [torch.from_numpy(f.copy()).cuda(), torch.from_numpy(g.copy()).cuda()]
    c.convolve_rows(td, 8, L)
    torch.cuda.synchronize()
    got = td[0].cpu().numpy()
    err = max(O.rel_l2(got[i], O.conv_complex(f[i], g[i])) for i in range(8))
    report("cfg5 4096 x 1-D complex L=8192", ms, 3 * rows * L * 16, err, [c.params(0)])
    # batch sweep of the same rows
    for nb in (1, 16, 256, 1024):
        ms = timeit(lambda: c.convolve_rows(d, nb, L, normalized=False), 10)
        print(json.dumps({"config": "cfg5 batch sweep", "rows": nb, "ms": ms,
                          "rows_per_s": nb / (ms / 1e3)}))
    del d, td
    # cfg5 layout (ii): the same 4096 signals interleaved ("Many": C=S=4096) --
    # forward of both inputs, multiply on the transformed data, backward; the
    # fused kernels need contiguous rows, so this layout runs unfused
    P = fp.Pad(0, L, 2 * L, rows, rows, 0, 0, -1, A=2, B=1)
    f = torch.zeros((L, rows), dtype=torch.complex128, device="cuda")
    F = torch.zeros(P.outputSize, dtype=torch.complex128, device="cuda")
    calls = P.residue_calls()

    def many_roundtrip():
        for r in calls:
            P.forward(f, r, F)
            P.backward(F, f, r)
    ms = timeit(many_roundtrip, 5)
    print(json.dumps({"config": "cfg5 layout (ii): fftPad(8192,16384,C=S=4096) forward+backward, "
                                "all residues", "ms": ms, "m": P.m, "p": P.p, "q": P.q, "D": P.D,
                      "GBps": (1 + 2 * P.q) * rows * L * 16 / 1e9 / (ms / 1e3)}))
    P.close()


if __name__ == "__main__":
    main()
