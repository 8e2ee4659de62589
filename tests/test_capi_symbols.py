"""CPU: lib_fftwpp.so loads and exports every symbol declared in include/*.h."""
import ctypes
import os
import re

import fftwpp_b200 as fp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef[^;]*;", "", text)  # callback types are not symbols
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_gpu_abi_symbols_exported():
    names = declared("fftwpp_gpu.h")
    assert len(names) >= 20
    lib = ctypes.CDLL(fp.lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_wrapper_api_symbols_exported():
    names = declared("cfftwpp.h")
    assert "fftwpp_create_conv1d" in names and "fftwpp_hconv3d_convolve" in names
    lib = ctypes.CDLL(fp.lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_reference_wrapper_symbol_set():
    # every entry point reference wrappers/cfftw++.cc:27-163 defines
    ref_defined = """create_doubleAlign delete_doubleAlign create_complexAlign
    delete_complexAlign get_fftwpp_maxthreads set_fftwpp_maxthreads
    fftwpp_create_conv1d fftwpp_conv1d_delete fftwpp_conv1d_convolve
    fftwpp_create_hconv1d fftwpp_hconv1d_delete fftwpp_HermitianSymmetrize
    fftwpp_hconv1d_convolve fftwpp_create_conv2d fftwpp_conv2d_delete
    fftwpp_HermitianSymmetrizeX fftwpp_conv2d_convolve fftwpp_create_hconv2d
    fftwpp_hconv2d_delete fftwpp_hconv2d_convolve fftwpp_create_conv3d
    fftwpp_conv3d_delete fftwpp_HermitianSymmetrizeXY fftwpp_conv3d_convolve
    fftwpp_create_hconv3d fftwpp_hconv3d_delete fftwpp_hconv3d_convolve""".split()
    lib = ctypes.CDLL(fp.lib_path)
    assert not [n for n in ref_defined if not hasattr(lib, n)]
