"""CPU: the host C++ classes (lib_fftwpp.so) must report the reference's own
bookkeeping -- every size accessor, the residue-call sequence and the full
index(r,i) table -- for a sweep of (kind,L,M,m,C,S,D,I).  Golden values come
from the reference itself (tests/golden/pad_params.json).  No GPU needed: GPU
plans are created lazily."""
import json
import os

import pytest

import fftwpp_b200 as fp

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "pad_params.json")) as fh:
    RECS = json.load(fh)


def test_golden_sweep_is_substantial():
    assert len(RECS) > 500
    assert {r["args"][0] for r in RECS} == {0, 1, 2, 3}


@pytest.mark.parametrize("chunk", range(8))
def test_accessors_and_index_tables(chunk):
    for rec in RECS[chunk::8]:
        kind, L, M, C, S, m, D, I = rec["args"]
        pad = fp.Pad(kind, L, M, C, S, m, D, I, A=2, B=1)
        for k, v in rec["info"].items():
            if k == "l" and kind == 2:
                continue  # uninitialised in the reference for fftPadHermitian
            assert pad.info[k] == v, (rec["args"], k, pad.info[k], v)
        assert pad.overwrite == 0
        calls = pad.residue_calls()
        assert calls == [c["r"] for c in rec["calls"]], rec["args"]
        for c in rec["calls"]:
            r = c["r"]
            assert pad.increment(r) == c["increment"]
            assert pad.blocksize(r) == c["blocksize"]
            assert pad.noutputs(r) == c["noutputs"]
            assert pad.span(r) == c["span"]
            got = [pad.index(r, i) for i in range(len(c["index"]))]
            assert got == c["index"], (rec["args"], r)
        pad.close()


def test_chooser_picks_valid_power_of_two_hybrid():
    pad = fp.Pad(fp.KIND_COMPLEX, 512, 1024, A=2, B=1, mult=fp.MULT_BINARY)
    assert (pad.m, pad.p, pad.q) == (512, 1, 2)
    pad = fp.Pad(fp.KIND_REAL, 512, 1024, C=512 * 512, A=2, B=1)
    assert (pad.m, pad.q) == (512, 2)
    pad = fp.Pad(fp.KIND_CENTERED, 256, 384, C=128, A=2, B=1)
    assert pad.p % 2 == 0 or pad.q == 1
    pad = fp.Pad(fp.KIND_HERMITIAN, 256, 384, A=2, B=1, mult=fp.MULT_REALBINARY)
    assert (pad.p == 2 and pad.D == 2) or pad.q == 1


def test_no_cpu_fallback():
    """Without a CUDA device compute entry points must fail loudly."""
    import subprocess
    import sys
    if fp.lib.fftwpp_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    code = ("import sys; sys.path.insert(0, %r); import numpy as np; import fftwpp_b200 as fp;"
            "c=fp.HybridConv([8],[16]); a=[np.zeros(8,complex),np.zeros(8,complex)];"
            "c.convolve(a); print('SILENT')" % os.path.dirname(HERE))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode != 0
    assert "SILENT" not in p.stdout
    assert "no CUDA device" in p.stderr or "failed" in p.stderr
