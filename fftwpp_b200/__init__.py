"""fftwpp_b200 -- B200-native hybrid dealiased convolutions (FFTW++ convolve.h path).

The product is the in-tree shared library ``lib_fftwpp.so`` (hand-written
sm_100a CUDA kernels + a thin C ABI + host C++ classes mirroring the
reference's convolve.h).  This package is only the ctypes binding to it
(mirroring reference wrappers/fftwpp.py) plus the distributed driver.  There is
no CPU compute path: importing works anywhere, computing needs a CUDA device.
"""
from ._lib import lib, lib_path, LibraryMissing  # noqa: F401
from .api import (  # noqa: F401
    Pad, HybridConv, KIND_COMPLEX, KIND_CENTERED, KIND_HERMITIAN, KIND_REAL,
    MULT_NONE, MULT_BINARY, MULT_REALBINARY, MULT_CORRELATION,
    FAMILY_COMPLEX, FAMILY_HERMITIAN, FAMILY_REAL, launch_count, set_stream,
    profile_enable, profile_read, pinned_array,
)
