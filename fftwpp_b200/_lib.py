"""ctypes loader for the in-tree lib_fftwpp.so (built by fftwpp_b200/Makefile)."""
import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
# FFTWPP_LIB selects another in-tree build (kernel A/B experiments)
lib_path = os.path.join(_here, os.environ.get("FFTWPP_LIB", "lib_fftwpp.so"))


class LibraryMissing(ImportError):
    pass


def _load():
    if not os.path.exists(lib_path):
        raise LibraryMissing(
            "%s is missing: run `make -C fftwpp_b200` (or __graft_entry__.build()); "
            "there is no Python/CPU fallback" % lib_path)
    return ctypes.CDLL(lib_path)


lib = _load()

c_size_t = ctypes.c_size_t
c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_long = ctypes.c_long
c_double = ctypes.c_double
c_u64 = ctypes.c_uint64
P = ctypes.POINTER


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


# generic handle API (include/cfftwpp.h part 2)
_sig("fftwpp_pad_create", c_void_p, c_int, c_size_t, c_size_t, c_size_t, c_size_t,
     c_size_t, c_size_t, c_long, c_size_t, c_size_t, c_int)
_sig("fftwpp_pad_destroy", None, c_void_p)
_sig("fftwpp_pad_info", None, c_void_p, P(c_size_t))
for _n in ("increment", "blocksize", "noutputs", "span"):
    _sig("fftwpp_pad_" + _n, c_size_t, c_void_p, c_size_t)
_sig("fftwpp_pad_index", c_size_t, c_void_p, c_size_t, c_size_t)
_sig("fftwpp_pad_forward", None, c_void_p, c_void_p, c_void_p, c_size_t)
_sig("fftwpp_pad_backward", None, c_void_p, c_void_p, c_void_p, c_size_t)
_sig("fftwpp_pad_forward_split", c_int, c_void_p, c_void_p, c_void_p, c_size_t)
_sig("fftwpp_pad_backward_split", c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_double)
_sig("fftwpp_conv_create", c_void_p, c_int, c_int, P(c_size_t), P(c_size_t),
     P(c_size_t), P(c_size_t), P(c_long), c_size_t, c_size_t, c_size_t, c_size_t, c_int)
_sig("fftwpp_conv_destroy", None, c_void_p)
_sig("fftwpp_conv_params", None, c_void_p, c_int, P(c_size_t))
_sig("fftwpp_conv_doubles", c_size_t, c_void_p)
_sig("fftwpp_conv_convolve", None, c_void_p, P(c_void_p), c_int)
_sig("fftwpp_conv_set_plane_chunk", None, c_void_p, c_size_t)
_sig("fftwpp_set_stream", None, c_void_p)

# reference wrapper API (include/cfftwpp.h part 1; reference wrappers/cfftw++.cc)
for _n in ("conv1d", "hconv1d"):
    _sig("fftwpp_create_" + _n, c_void_p, c_size_t)
for _n in ("conv2d", "hconv2d"):
    _sig("fftwpp_create_" + _n, c_void_p, c_size_t, c_size_t)
for _n in ("conv3d", "hconv3d"):
    _sig("fftwpp_create_" + _n, c_void_p, c_size_t, c_size_t, c_size_t)
for _n in ("conv1d", "hconv1d", "conv2d", "hconv2d", "conv3d", "hconv3d"):
    _sig("fftwpp_%s_delete" % _n, None, c_void_p)
    _sig("fftwpp_%s_convolve" % _n, None, c_void_p, c_void_p, c_void_p)
_sig("fftwpp_HermitianSymmetrize", None, c_void_p)
_sig("fftwpp_HermitianSymmetrizeX", None, c_size_t, c_size_t, c_size_t, c_void_p)
_sig("fftwpp_HermitianSymmetrizeXY", None, c_size_t, c_size_t, c_size_t, c_size_t,
     c_size_t, c_void_p)
_sig("get_fftwpp_maxthreads", c_size_t)
_sig("set_fftwpp_maxthreads", None, c_size_t)

# thin GPU C ABI (include/fftwpp_gpu.h)
_sig("fftwpp_gpu_device_count", c_int)
_sig("fftwpp_gpu_launch_count", c_u64)
_sig("fftwpp_gpu_last_error", ctypes.c_char_p)
_sig("fftwpp_gpu_device_sync", c_int)
_sig("fftwpp_gpu_set_device", c_int, c_int)
_sig("fftwpp_gpu_profile_enable", c_int, c_int)
_sig("fftwpp_gpu_profile_read", c_int, P(c_double), P(c_u64))
_sig("fftwpp_gpu_malloc_host", c_int, P(c_void_p), c_size_t)
_sig("fftwpp_gpu_free_host", c_int, c_void_p)
_sig("fftwpp_gpu_comm_unique_id", c_int, ctypes.c_char_p)
_sig("fftwpp_gpu_comm_create", c_int, c_int, c_int, ctypes.c_char_p, P(c_void_p))
_sig("fftwpp_gpu_comm_destroy", c_int, c_void_p)
_sig("fftwpp_mpiconv3_create", c_void_p, c_int, P(c_size_t), P(c_size_t), P(c_size_t),
     P(c_size_t), P(c_long), c_size_t, c_size_t, c_int, c_int, c_int, c_void_p)
_sig("fftwpp_mpiconv3_create_pencil", c_void_p, c_int, P(c_size_t), P(c_size_t), P(c_size_t),
     P(c_size_t), P(c_long), c_size_t, c_size_t, c_int, c_int, c_int, c_void_p, c_int, c_int,
     c_void_p)
_sig("fftwpp_mpiconv3_destroy", None, c_void_p)
_sig("fftwpp_mpiconv3_split", None, c_void_p, P(c_size_t))
_sig("fftwpp_mpiconv3_params", None, c_void_p, c_int, P(c_size_t))
_sig("fftwpp_mpiconv3_convolve", None, c_void_p, P(c_void_p), c_int)
_sig("fftwpp_mpiconv3_convolve_async", None, c_void_p, P(c_void_p), c_int, c_int)
_sig("fftwpp_mpiconv3_wait", None, c_void_p, c_int)
_sig("fftwpp_mpiconv3_exchange_table", None, c_void_p, c_int, P(ctypes.c_ulonglong),
     P(ctypes.c_ulonglong), P(ctypes.c_ulonglong), P(ctypes.c_ulonglong))
_sig("fftwpp_mpiconv3_set_plane_chunk", None, c_void_p, c_size_t)
_sig("fftwpp_mpiconv3_symmetrize", None, c_void_p, c_void_p)
_sig("fftwpp_mpiconv2_create", c_void_p, c_int, P(c_size_t), P(c_size_t), P(c_size_t),
     P(c_size_t), P(c_long), c_size_t, c_size_t, c_int, c_int, c_int, c_void_p)
_sig("fftwpp_mpiconv2_destroy", None, c_void_p)
_sig("fftwpp_mpiconv2_split", None, c_void_p, P(c_size_t))
_sig("fftwpp_mpiconv2_params", None, c_void_p, c_int, P(c_size_t))
_sig("fftwpp_mpiconv2_convolve", None, c_void_p, P(c_void_p), c_int)
_sig("fftwpp_mpiconv2_exchange_table", None, c_void_p, c_int, P(ctypes.c_ulonglong),
     P(ctypes.c_ulonglong), P(ctypes.c_ulonglong), P(ctypes.c_ulonglong))
_sig("fftwpp_mpifft_create", c_void_p, c_int, c_int, P(c_size_t), c_int, c_int, c_int, c_void_p)
_sig("fftwpp_mpifft_destroy", None, c_void_p)
_sig("fftwpp_mpifft_split", None, c_void_p, P(c_size_t))
_sig("fftwpp_mpifft_words", c_size_t, c_void_p)
_sig("fftwpp_mpifft_exchange_table", None, c_void_p, c_int, P(ctypes.c_ulonglong),
     P(ctypes.c_ulonglong), P(ctypes.c_ulonglong), P(ctypes.c_ulonglong))
_sig("fftwpp_mpifft_forward", None, c_void_p, c_void_p, c_void_p)
_sig("fftwpp_mpifft_backward", None, c_void_p, c_void_p, c_void_p)
_sig("fftwpp_mpifft_normalize", None, c_void_p, c_void_p)
_sig("fftwpp_mpifft_shift", None, c_void_p, c_void_p)
_sig("fftwpp_mpifft_denyquist", None, c_void_p, c_void_p)
HOST_MULT = ctypes.CFUNCTYPE(None, P(c_void_p), c_size_t, c_void_p, c_size_t)
DEVICE_MULT = ctypes.CFUNCTYPE(None, P(c_void_p), c_size_t, c_void_p, c_void_p)
_sig("fftwpp_conv_create_custom", c_void_p, c_int, c_int, P(c_size_t), P(c_size_t),
     P(c_size_t), P(c_size_t), P(c_long), c_size_t, c_size_t, c_size_t, c_size_t,
     HOST_MULT, DEVICE_MULT)
_sig("fftwpp_indices_get", None, c_void_p, P(c_size_t), P(c_size_t))
_sig("fftwpp_indices_size", c_size_t, c_void_p)
_sig("fftwpp_indices_outer", c_size_t, c_void_p, c_size_t)
_sig("fftwpp_indices_index", c_size_t, c_void_p, c_size_t)
_sig("fftwpp_conv_convolve_async", None, c_void_p, P(c_void_p), c_int, c_int)
_sig("fftwpp_conv_wait", None, c_void_p, c_int)
_sig("fftwpp_conv_convolve_rows", None, c_void_p, P(c_void_p), c_size_t, c_size_t, c_int)
