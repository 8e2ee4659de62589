"""Source compatibility of the C++ boundary: the reference's own
callers -- examples/exampleconv*.cc, tests/hybrid*.cc with tests/options.cc and
tests/direct.cc, wrappers/cexample.c -- compile UNMODIFIED against
fftwpp_b200/cpp/*.h and link with lib_fftwpp.so (tests/refprogs/Makefile).
CPU-side check: needs the reference sources, so it is skipped where
/root/reference is absent (the GPU box runs the prebuilt binaries instead,
tests/test_gpu_refprogs.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIR = os.path.join(ROOT, "tests", "refprogs")
PROGS = ["exampleconv", "exampleconv2", "exampleconv3", "exampleconvh", "exampleconvh2",
         "exampleconvh3", "exampleconvr", "exampleconvr2", "exampleconvr3", "cexample",
         "hybridconv", "hybridconv2", "hybridconv3", "hybridconvh", "hybridconvh2",
         "hybridconvh3", "hybridconvr", "hybridconvr2", "hybridconvr3", "hybrid", "hybridh",
         "hybridr", "cexample_refwrap"]
WRAPPER_SYMBOLS = ["create_doubleAlign", "delete_doubleAlign", "create_complexAlign",
                   "delete_complexAlign", "get_fftwpp_maxthreads", "set_fftwpp_maxthreads",
                   "fftwpp_HermitianSymmetrize", "fftwpp_HermitianSymmetrizeX",
                   "fftwpp_HermitianSymmetrizeXY"] + \
    ["fftwpp_create_%s" % k for k in ("conv1d", "hconv1d", "conv2d", "hconv2d", "conv3d", "hconv3d")] + \
    ["fftwpp_%s_%s" % (k, op) for k in ("conv1d", "hconv1d", "conv2d", "hconv2d", "conv3d", "hconv3d")
     for op in ("convolve", "delete")]


@pytest.mark.skipif(not os.path.exists("/root/reference/convolve.h"),
                    reason="reference sources not present on this host")
def test_reference_callers_compile_unmodified():
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    r = subprocess.run(["make", "-j8", "-C", DIR], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for p in PROGS:
        assert os.access(os.path.join(DIR, "_build", p), os.X_OK), p


@pytest.mark.skipif(not os.path.exists("/root/reference/convolve.h"),
                    reason="reference sources not present on this host")
def test_reference_wrapper_library_compiles_unmodified():
    """wrappers/cfftw++.cc itself (the reference's C wrapper LIBRARY, not just its
    callers) builds against cpp/HybridConvolution.h, convolve.h, Complex.h and
    cfftw++.h, and exports the 27 symbols of SURVEY 8(b); the C example linked on
    top of it resolves the wrapper API to that library."""
    r = subprocess.run(["make", "-C", DIR, "_build/libcfftw_refwrap.so", "_build/cexample_refwrap"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lib = os.path.join(DIR, "_build", "libcfftw_refwrap.so")
    nm = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    have = {l.split()[-1] for l in nm.splitlines() if " T " in l}
    assert set(WRAPPER_SYMBOLS) <= have, sorted(set(WRAPPER_SYMBOLS) - have)
    ldd = subprocess.run(["ldd", os.path.join(DIR, "_build", "cexample_refwrap")],
                         capture_output=True, text=True).stdout
    assert "libcfftw_refwrap.so" in ldd and "lib_fftwpp.so" in ldd
    assert ldd.index("libcfftw_refwrap.so") < ldd.index("lib_fftwpp.so")


@pytest.mark.skipif(not os.path.exists(os.path.join(DIR, "_build", "hybridconv")),
                    reason="tests/refprogs/_build not built")
def test_reference_caller_fails_loudly_without_gpu():
    """No CPU fallback: on a host without a CUDA device the unmodified
    reference driver must exit non-zero with the library's message."""
    import fftwpp_b200 as fp
    if fp.lib.fftwpp_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([os.path.join(DIR, "_build", "hybridconv"), "-L8", "-M16", "-E", "-T1"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "fftwpp-b200:" in r.stderr  # the library's own error line (cerr + exit)


@pytest.mark.skipif(not os.path.exists("/root/reference/wrappers/fftwpp.py"),
                    reason="reference sources not present on this host")
def test_reference_python_wrapper_binds_unmodified(tmp_path):
    """The reference's own wrappers/fftwpp.py loads `lib_fftwpp.so` from the
    directory it sits in (fftwpp.py:24-26) and sets a prototype on every wrapper
    symbol at import time.  Placed (here: symlinked, nothing is copied) next to
    this repository's library, it must import: every symbol it names exists."""
    (tmp_path / "fftwpp.py").symlink_to("/root/reference/wrappers/fftwpp.py")
    (tmp_path / "lib_fftwpp.so").symlink_to(os.path.join(ROOT, "fftwpp_b200", "lib_fftwpp.so"))
    code = ("import sys; sys.path.insert(0, %r); import fftwpp; "
            "assert fftwpp.base == %r, fftwpp.base; "
            "print(sorted(fftwpp.__all__), fftwpp.fftwpp_get_maxthreads() >= 1)") % (str(tmp_path), str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "['Convolution', 'HConvolution'] True" in r.stdout
