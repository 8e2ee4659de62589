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


def test_chooser_baseline_configs():
    """The deterministic chooser's picks for the BASELINE configs (pinned so a
    cost-model edit that changes the measured configurations is noticed)."""
    def params(c, d):
        p = c.params(d)
        return p["m"], p["p"], p["q"]
    c = fp.HybridConv([1 << 20], [1 << 21])       # cfg1: two-stage inner, stage B on long rows
    assert params(c, 0) == (4096, 256, 512)
    c = fp.HybridConv([1 << 22], [1 << 23])
    assert params(c, 0) == (8192, 512, 1024)
    c = fp.HybridConv([1 << 20], [1 << 21], A=3, B=1, mult=fp.MULT_NONE)   # no long-row stage B
    assert params(c, 0)[0] <= 4096
    c = fp.HybridConv([4096, 4096], [8192, 8192])                 # cfg2: fused rows at m=4096
    assert params(c, 0) == (4096, 1, 2) and params(c, 1) == (4096, 1, 2)
    c = fp.HybridConv([256] * 3, [384] * 3, family=fp.FAMILY_HERMITIAN)   # cfg3
    assert [params(c, d) for d in range(3)] == [(128, 2, 3)] * 3
    c = fp.HybridConv([512] * 3, [1024] * 3, family=fp.FAMILY_REAL)       # cfg4 (headline)
    assert [params(c, d) for d in range(3)] == [(512, 1, 2)] * 3
    c = fp.HybridConv([8192], [16384])                # cfg5: fused long rows (tensor memory)
    assert params(c, 0) == (8192, 1, 2)


def test_partially_forced_parameters():
    """m alone (D = 0, I = -1), as Application(A,B,mult,threads,verbose,m) in the
    reference (convolve.h:95-123): m is honoured, D is chosen.  (Used to divide
    by zero in the forced constructor.)"""
    for kind_args in (dict(), dict(family=fp.FAMILY_HERMITIAN), dict(family=fp.FAMILY_REAL)):
        c = fp.HybridConv([512], [1000], m=[256], **kind_args)
        p = c.params(0)
        assert p["m"] == 256 and p["D"] >= 1 and p["m"] * p["q"] >= 1000
        c.close()
    c = fp.HybridConv([5000], [10000], m=[8192])
    assert (c.params(0)["m"], c.params(0)["p"], c.params(0)["q"]) == (8192, 1, 2)
    c.close()
    os.environ["FFTWPP_NO_LONG_ROWS"] = "1"       # the two-stage path for long rows
    try:
        c = fp.HybridConv([8192], [16384])
        assert c.params(0)["p"] > 2
        c.close()
    finally:
        del os.environ["FFTWPP_NO_LONG_ROWS"]


def test_device_multiplier_needs_host_function():
    """fftwpp_conv_create_custom identifies the device multiplier by the host
    function's address, so a NULL host function is refused."""
    import subprocess
    import sys
    code = ("import ctypes, fftwpp_b200 as fp\n"
            "from fftwpp_b200._lib import lib, HOST_MULT, DEVICE_MULT\n"
            "a = (ctypes.c_size_t * 1)(8); b = (ctypes.c_size_t * 1)(16)\n"
            "lib.fftwpp_conv_create_custom(1, 0, a, b, None, None, None, 0, 0, 2, 1,\n"
            "    ctypes.cast(None, HOST_MULT), ctypes.cast(None, DEVICE_MULT))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       cwd=os.path.dirname(HERE))
    assert r.returncode != 0
    assert "host multiplier is required" in r.stderr


@pytest.mark.parametrize("args,needle", [
    ((1, 9, 27, 1, 1, 3, 1, 0), "Invalid parameters"),     # centred with odd p
    ((0, 8, 32, 1, 1, 8, 3, 0), "Invalid parameters"),     # odd D < n
    ((3, 12, 24, 1, 1, 4, 1, 0), "Invalid parameters"),    # real: n even with odd p > 2
    ((0, 16, 8, 1, 1, 0, 0, -1), "is greater than M"),     # L > M
])
def test_error_policy_matches_reference(args, needle):
    """Errors are a message on stderr and exit(-1), never an exception or a
    return code (reference convolve.h:226-232, convolve.cc:480-501)."""
    import subprocess
    import sys
    code = "import fftwpp_b200 as fp\nfp.Pad(*%r)\nprint('constructed')\n" % (args,)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       cwd=os.path.dirname(HERE))
    assert r.returncode != 0 and "constructed" not in r.stdout
    assert needle in r.stderr


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


def test_cuda_array_interface_pointers():
    """device arrays of other libraries (cupy, numba) reach the C ABI through
    __cuda_array_interface__; only the pointer extraction is
    checked here, no device is touched"""
    from fftwpp_b200.api import _ptr

    class Dev:
        def __init__(self, shape, strides=None):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<c16", "version": 3,
                                             "data": (0x7f0000001000, False), "strides": strides}
    assert _ptr(Dev((4, 8))) == 0x7f0000001000
    assert _ptr(Dev((4, 8), strides=(128, 16))) == 0x7f0000001000
    with pytest.raises(ValueError):
        _ptr(Dev((4, 8), strides=(256, 16)))
    with pytest.raises(TypeError):
        _ptr([1, 2, 3])
