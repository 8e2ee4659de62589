"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF
(oracle/_ref: the unmodified /root/reference sources built by oracle/Makefile
against the FFTW3-API shim) in the build container.  The reference ships no
golden vectors of its own (every expected value in its tests is computed at
test time), so these fixtures are what pins the oracle and the
host bookkeeping on machines where /root/reference is absent.

  python tests/golden/make_golden.py

Writes
  pad_params.json   for a sweep of (kind,L,M,m,C,S,D,I): every size accessor,
                    the residue-call sequence, increment/blocksize/noutputs/span
                    per call and the full index(r,i) table
  conv_cases.npz    seeded inputs and the reference's hybrid-convolution outputs
                    (and its FFT-free direct results) for 1/2/3-D complex, real
                    and Hermitian cases incl. odd sizes and explicit (m,D,I)
  forward_cases.npz reference fft->forward(f,F,r) outputs for a few residue
                    layouts (incl. conjugate-pair D>1 and real packed residues)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402

KEYS = ("L M C S m p q n R dr D D0 l b inplace centered inputLength wordSize doubles "
        "outputSize workSizeW paddedSize normalization repad conjugates residueBlocks").split()


def valid_D(kind, m, p, q, n, S, C):
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


def pad_sweep():
    recs = []
    for kind in (0, 1, 2, 3):
        for L in (3, 4, 5, 8, 12):
            for M in sorted(set([L, 2 * L, (3 * L + 1) // 2, 5 * L // 2, 4 * L])):
                for m in sorted(set([M, L + 1, L, (L + 1) // 2, max(2, L // 4)])):
                    for (C, S) in ((1, 1), (2, 3)):
                        if kind == 2 and S != C:
                            continue
                        p, n, q = O.parameters(L, M, m, kind in (1, 2))
                        if q * m < M:
                            continue
                        if kind == 1 and q > 1 and p % 2:
                            continue
                        for D in valid_D(kind, m, p, q, n, S, C):
                            I = 0
                            b = R.RefPad(kind, L, M, C, S, m, D, I, A=2, B=1)
                            if b.overwrite:
                                b.close()
                                continue
                            rec = {"args": [kind, L, M, C, S, m, D, I],
                                   "info": {k: b.info[k] for k in KEYS}, "calls": []}
                            for r in b.residue_calls():
                                no = b.noutputs(r)
                                cnt = no * ((b.D0 if r == 0 else b.D) if kind == 2 else 1)
                                rec["calls"].append({
                                    "r": r, "increment": b.increment(r),
                                    "blocksize": b.blocksize(r), "noutputs": no,
                                    "span": b.span(r),
                                    "index": [b.index(r, i) for i in range(cnt)]})
                            recs.append(rec)
                            b.close()
    return recs


def conv_cases():
    rng = np.random.default_rng(20261017)
    out = {}
    cases = [
        # name, family, L, M, m, D, I
        ("c1_auto8", 0, [8], [16], None, None, None),
        ("c1_m4D2", 0, [8], [16], [4], [2], [0]),
        ("c1_odd", 0, [7], [20], [5], [1], [0]),
        ("c1_inner", 0, [12], [30], [4], [1], [0]),
        ("c2", 0, [5, 6], [10, 12], None, None, None),
        ("c3", 0, [4, 3, 5], [8, 6, 10], None, None, None),
        ("r1", 2, [9], [18], None, None, None),
        ("r1_m8", 2, [8], [16], [8], [1], [0]),
        ("r2", 2, [6, 5], [12, 10], None, None, None),
        ("r3", 2, [4, 5, 6], [8, 10, 12], None, None, None),
        ("r3_forced", 2, [8, 8, 8], [16, 16, 16], [8, 4, 8], [1, 1, 1], [0, 0, 0]),
        ("h1_even", 1, [8], [12], None, None, None),
        ("h1_odd", 1, [7], [11], None, None, None),
        ("h2", 1, [8, 6], [12, 9], None, None, None),
        ("h2_odd", 1, [7, 5], [11, 8], None, None, None),
        ("h3", 1, [6, 5, 8], [9, 8, 12], None, None, None),
        # inner (p > 2) routines of the real and Hermitian classes
        ("r1_inner_even", 2, [16], [32], [4], [1], [0]),
        ("r1_inner_odd", 2, [12], [36], [4], [1], [0]),
        ("r2_inner", 2, [12, 6], [36, 12], [4, 6], [1, 1], [0, 0]),
        ("h1_inner", 1, [16], [24], [4], [2], [0]),
        ("h2_inner", 1, [8, 16], [12, 24], [4, 4], [1, 2], [0, 0]),
    ]
    meta = []
    for name, fam, L, M, m, D, I in cases:
        if fam == 2:
            f, g = rng.uniform(-1, 1, L), rng.uniform(-1, 1, L)
        else:
            shp = L if fam == 0 else L[:-1] + [(L[-1] + 1) // 2]
            f = rng.uniform(-1, 1, shp) + 1j * rng.uniform(-1, 1, shp)
            g = rng.uniform(-1, 1, shp) + 1j * rng.uniform(-1, 1, shp)
            if fam == 1:
                R.symmetrize(L, f)
                R.symmetrize(L, g)
        c = R.RefConv(L, M, family=fam, m=m, D=D, I=I, threads=1)
        a = [np.ascontiguousarray(f.copy()), np.ascontiguousarray(g.copy())]
        c.convolve(a)
        params = [c.params(d) for d in range(len(L))]
        c.close()
        if fam == 0:
            direct = R.direct_complex(f, g)
        elif fam == 2:
            direct = R.direct_real(f, g)
        else:
            direct = R.direct_hermitian(L, f, g)
        out[name + "_f"] = f
        out[name + "_g"] = g
        out[name + "_hybrid"] = a[0]
        out[name + "_direct"] = direct
        meta.append({"name": name, "family": fam, "L": L, "M": M, "m": m, "D": D, "I": I,
                     "params": params})
    # centred 1-D complex direct convolution (tests/direct.h:27-39)
    for L in (8, 7):
        f = rng.uniform(-1, 1, L) + 1j * rng.uniform(-1, 1, L)
        g = rng.uniform(-1, 1, L) + 1j * rng.uniform(-1, 1, L)
        out["cen%d_f" % L] = f
        out["cen%d_g" % L] = g
        out["cen%d_direct" % L] = R.direct_centered1(f, g)
    return out, meta


def forward_cases():
    rng = np.random.default_rng(777)
    out = {}
    meta = []
    cases = [
        ("std_p1_D2", 0, 8, 24, 1, 1, 8, 2),
        ("std_p2_D1", 0, 8, 16, 1, 1, 4, 1),
        ("std_p2_D4", 0, 8, 16, 1, 1, 4, 4),
        ("std_inner", 0, 12, 36, 1, 1, 4, 1),
        ("std_many", 0, 5, 15, 2, 3, 5, 1),
        ("cen_p2_D2", 1, 8, 16, 1, 1, 4, 2),
        ("cen_many", 1, 8, 12, 2, 2, 4, 1),
        ("herm_p2", 2, 8, 12, 1, 1, 4, 2),
        ("herm_many", 2, 8, 16, 2, 2, 4, 2),
        ("real_p1", 3, 8, 32, 1, 1, 8, 1),
        ("real_p1_D", 3, 8, 40, 1, 1, 8, 2),
        ("real_p2", 3, 8, 16, 1, 1, 4, 1),
        ("real_many", 3, 6, 12, 2, 3, 6, 1),
        ("herm_inner", 2, 16, 24, 1, 1, 4, 2),
        ("real_inner_even", 3, 16, 32, 1, 1, 4, 1),
        ("real_inner_odd", 3, 12, 36, 1, 1, 4, 1),
        ("real_inner_D2", 3, 12, 60, 1, 1, 4, 2),
        ("real_inner_many", 3, 12, 36, 2, 3, 4, 1),
    ]
    for name, kind, L, M, C, S, m, D in cases:
        b = R.RefPad(kind, L, M, C, S, m, D, 0, A=1, B=1)
        Lin = (L + 1) // 2 if kind == 2 else L
        if kind == 3:
            f = np.zeros((Lin, S))
            f[:, :C] = rng.uniform(-1, 1, (Lin, C))
        else:
            f = np.zeros((Lin, S), dtype=np.complex128)
            f[:, :C] = rng.uniform(-1, 1, (Lin, C)) + 1j * rng.uniform(-1, 1, (Lin, C))
            if kind == 2:
                f[0] = f[0].real
        out[name + "_f"] = f
        calls = b.residue_calls()
        for r in calls:
            out["%s_F%d" % (name, r)] = b.forward(f.copy(), r)
        meta.append({"name": name, "args": [kind, L, M, C, S, m, D, 0], "calls": calls})
        b.close()
    return out, meta


def main():
    recs = pad_sweep()
    with open(os.path.join(HERE, "pad_params.json"), "w") as fh:
        json.dump(recs, fh, separators=(",", ":"))
    conv, meta = conv_cases()
    np.savez_compressed(os.path.join(HERE, "conv_cases.npz"), **conv)
    fwd, fmeta = forward_cases()
    np.savez_compressed(os.path.join(HERE, "forward_cases.npz"), **fwd)
    with open(os.path.join(HERE, "cases_meta.json"), "w") as fh:
        json.dump({"conv": meta, "forward": fmeta}, fh, indent=1)
    print("pad records:", len(recs), "conv cases:", len(meta), "forward cases:", len(fmeta))


if __name__ == "__main__":
    main()
