"""Distributed (slab over y) 2-D and 3-D hybrid convolutions: Python handles on
the C++ Convolution2MPI / Convolution3MPI counterparts (cpp/mpiconvolve.h;
reference mpi/mpiconvolve.h:72-305, drivers mpi/tests/hybridconv2.cc,
hybridconvr3.cc, hybridconvh3.cc).  One process per GPU;
torch.distributed is used only to bootstrap the NCCL communicator (broadcast of
the unique id) -- the exchange itself is issued by lib_fftwpp.so."""
import ctypes

from ._lib import lib
from .api import _ptr, FAMILY_REAL, MULT_BINARY


def local_dimension(N, rank, size):
    """extent, start of rank's share (reference mpi/mpitranspose.h:118-130)."""
    n = (N + size - 1) // size
    s = n * rank
    if s >= N:
        return 0, N
    return (n if s + n <= N else N - s), s


def _nccl_comm(rank, world):
    """NCCL communicator of lib_fftwpp.so, bootstrapped over torch.distributed."""
    import torch
    import torch.distributed as dist
    comm = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        rc = lib.fftwpp_gpu_comm_unique_id(buf)
        if rc:
            raise RuntimeError(lib.fftwpp_gpu_last_error().decode())
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    rc = lib.fftwpp_gpu_comm_create(rank, world, raw, ctypes.byref(comm))
    if rc:
        raise RuntimeError(lib.fftwpp_gpu_last_error().decode())
    return comm


def _nccl_subcomm(rank, world, color, key):
    """NCCL communicator of the ranks sharing `color`, ordered by `key`
    (the MPI_Comm_split of reference mpi/mpigroup.h:46-48), bootstrapped over
    torch.distributed: the first rank of every colour creates a unique id,
    all ids are gathered, every rank joins its colour's communicator.
    Returns (comm, rank in group, group size)."""
    import torch
    import torch.distributed as dist
    info = torch.tensor([color, key], dtype=torch.int64, device="cuda")
    allinfo = [torch.zeros_like(info) for _ in range(world)]
    dist.all_gather(allinfo, info)
    members = sorted((int(t[1]), r) for r, t in enumerate(allinfo) if int(t[0]) == color)
    ranks = [r for _, r in members]
    me = ranks.index(rank)
    buf = ctypes.create_string_buffer(128)
    if me == 0:
        rc = lib.fftwpp_gpu_comm_unique_id(buf)
        if rc:
            raise RuntimeError(lib.fftwpp_gpu_last_error().decode())
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    ids = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(ids, t)
    raw = bytes(ids[ranks[0]].cpu().tolist())
    comm = ctypes.c_void_p()
    rc = lib.fftwpp_gpu_comm_create(me, len(ranks), raw, ctypes.byref(comm))
    if rc:
        raise RuntimeError(lib.fftwpp_gpu_last_error().decode())
    return comm, me, len(ranks)


class SlabConvolution2:
    """2-D complex convolution of Lx x Ly data, y split over the ranks
    (reference Convolution2MPI, mpi/mpiconvolve.h:72-179)."""

    def __init__(self, Lx, Ly, Mx, My, rank, world, m=None, D=None, I=None, A=2, B=1,
                 mult=MULT_BINARY, comm="nccl", family=0):
        """family 0: complex; family 1: centred Hermitian (x centred, y holds
        the ceil(Ly/2) non-negative modes, split over the ranks; reference
        mpi/tests/hybridconvh2.cc)."""
        self.family = family
        self.L, self.M = [Lx, Ly], [Mx, My]
        self.rank, self.world, self.A, self.B = rank, world, A, B
        self._comm = _nccl_comm(rank, world) if comm == "nccl" else ctypes.c_void_p()
        arr, larr = ctypes.c_size_t * 2, ctypes.c_long * 2
        m = arr(*([0] * 2 if m is None else m))
        D = arr(*([0] * 2 if D is None else D))
        I = larr(*([-1] * 2 if I is None else I))
        self._h = lib.fftwpp_mpiconv2_create(family, arr(*self.L), arr(*self.M), m, D, I,
                                             A, B, mult, rank, world, self._comm)
        buf = (ctypes.c_size_t * 9)()
        lib.fftwpp_mpiconv2_split(self._h, buf)
        self.split = dict(zip("X Y Z x y z x0 y0 z0".split(), [int(v) for v in buf]))

    def exchange_table(self, direction):
        n = self.world
        t = [(ctypes.c_ulonglong * n)() for _ in range(4)]
        lib.fftwpp_mpiconv2_exchange_table(self._h, direction, *t)
        return [[int(v) for v in a] for a in t]

    def local_shape(self):
        return (self.L[0], self.split["y"])

    def convolve(self, arrays, normalized=True):
        n = max(self.A, self.B)
        ptrs = (ctypes.c_void_p * n)(*[_ptr(a) for a in arrays[:n]])
        lib.fftwpp_mpiconv2_convolve(self._h, ptrs, 1 if normalized else 0)
        return arrays[0]

    def close(self):
        if self._h:
            lib.fftwpp_mpiconv2_destroy(self._h)
            self._h = None
        if self._comm:
            lib.fftwpp_gpu_comm_destroy(self._comm)
            self._comm = ctypes.c_void_p()


class SlabConvolution3:
    """family: FAMILY_COMPLEX, FAMILY_HERMITIAN (local slabs Lx x y x ceil(Lz/2)
    complex, centred; the caller symmetrises the global field as the
    reference's HermitianSymmetrizeXY(split3), mpi/mpiconvolve.cc:11-142) or
    FAMILY_REAL (doubles)."""

    def __init__(self, Lx, Ly, Lz, Mx, My, Mz, rank, world, family=FAMILY_REAL,
                 m=None, D=None, I=None, A=2, B=1, mult=MULT_BINARY, comm="nccl", grid=None):
        """grid=(py,pz) with py*pz == world selects the PENCIL decomposition
        (y split py ways, z split pz ways; rank = iy*pz + iz; reference
        mpi/mpigroup.h:33-50, mpi/mpiconvolve.h:208-216): the local arrays are
        Lx x y x z pencils.  Default: slabs over y."""
        self.L, self.M = [Lx, Ly, Lz], [Mx, My, Mz]
        self.rank, self.world, self.family, self.A, self.B = rank, world, family, A, B
        arr, larr = ctypes.c_size_t * 3, ctypes.c_long * 3
        m = arr(*([0] * 3 if m is None else m))
        D = arr(*([0] * 3 if D is None else D))
        I = larr(*([-1] * 3 if I is None else I))
        self.grid = None
        self._comm2 = ctypes.c_void_p()
        self.zsplit = {"z": Lz, "z0": 0}
        if grid is not None and grid[1] > 1:
            py, pz = grid
            if py * pz != world:
                raise ValueError("grid must multiply to the number of ranks")
            iy, iz = divmod(rank, pz)
            self.grid = (py, pz, iy, iz)
            self._comm, ry, sy = _nccl_subcomm(rank, world, color=iz, key=iy)   # same z slice
            self._comm2, rz, sz = _nccl_subcomm(rank, world, color=py + iy, key=iz)  # same y slice
            self._h = lib.fftwpp_mpiconv3_create_pencil(family, arr(*self.L), arr(*self.M), m, D, I,
                                                        A, B, mult, ry, sy, self._comm, rz, sz,
                                                        self._comm2)
            zl, z0 = local_dimension(Lz, rz, sz)
            self.zsplit = {"z": zl, "z0": z0}
        else:
            self._comm = _nccl_comm(rank, world) if comm == "nccl" else ctypes.c_void_p()
            self._h = lib.fftwpp_mpiconv3_create(family, arr(*self.L), arr(*self.M), m, D, I,
                                                 A, B, mult, rank, world, self._comm)
        buf = (ctypes.c_size_t * 9)()
        lib.fftwpp_mpiconv3_split(self._h, buf)
        self.split = dict(zip("X Y Z x y z x0 y0 z0".split(), [int(v) for v in buf]))

    def params(self):
        out = []
        for d in range(3):
            buf = (ctypes.c_size_t * 8)()
            lib.fftwpp_mpiconv3_params(self._h, d, buf)
            out.append(dict(zip("m p q n D inplace C S".split(), [int(v) for v in buf])))
        return out

    def exchange_table(self, direction):
        n = self.world
        t = [(ctypes.c_ulonglong * n)() for _ in range(4)]
        lib.fftwpp_mpiconv3_exchange_table(self._h, direction, *t)
        return [[int(v) for v in a] for a in t]

    def local_shape(self):
        return (self.L[0], self.split["y"], self.zsplit["z"])

    def full_inputs(self, seed=1234, scale_second=None):
        """The globally defined seeded fields (host numpy arrays, every rank
        generates the same ones); make_inputs() takes this rank's slab."""
        import numpy as np
        import torch
        Lx, Ly, Lz = self.L
        out = []
        for a in range(self.A):
            g = torch.Generator(device="cpu").manual_seed(seed + a)
            full = torch.rand((Lx, Ly, Lz), dtype=torch.float64, generator=g) * 2 - 1
            if a == 1:
                full = full * (scale_second if scale_second is not None
                               else 1.7 / np.sqrt(float(Lx * Ly * Lz)))
            if self.family == 0:
                g2 = torch.Generator(device="cpu").manual_seed(seed + 100 + a)
                im = torch.rand((Lx, Ly, Lz), dtype=torch.float64, generator=g2) * 2 - 1
                full = torch.complex(full, im * (full.abs().max()))
            out.append(full.numpy())
        return out

    def make_inputs(self, seed=1234, scale_second=None):
        """Seeded inputs: the local slab of a globally defined random field."""
        import torch
        y, y0 = self.split["y"], self.split["y0"]
        return [torch.from_numpy(a[:, y0:y0 + y, :].copy()).cuda()
                for a in self.full_inputs(seed, scale_second)]

    def convolve(self, arrays, normalized=True):
        n = max(self.A, self.B)
        ptrs = (ctypes.c_void_p * n)(*[_ptr(a) for a in arrays[:n]])
        lib.fftwpp_mpiconv3_convolve(self._h, ptrs, 1 if normalized else 0)
        return arrays[0]

    def convolve_raw(self, arrays):
        return self.convolve(arrays, normalized=False)

    def convolve_async(self, arrays, slot=0, normalized=True):
        """Pipelined form for PINNED host slabs (fftwpp_b200.pinned_array of
        local_shape()): returns at once; wait(slot) blocks until the result is
        back in arrays[0:B].  Collective, two slots."""
        n = max(self.A, self.B)
        ptrs = (ctypes.c_void_p * n)(*[_ptr(a) for a in arrays[:n]])
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[slot] = arrays
        lib.fftwpp_mpiconv3_convolve_async(self._h, ptrs, 1 if normalized else 0, slot)

    def wait(self, slot=0):
        lib.fftwpp_mpiconv3_wait(self._h, slot)
        if hasattr(self, "_inflight"):
            self._inflight.pop(slot, None)

    def set_plane_chunk(self, chunk):
        lib.fftwpp_mpiconv3_set_plane_chunk(self._h, int(chunk))

    def symmetrize(self, array):
        """Hermitian family: enforce the symmetry of the GLOBAL field on this
        rank's device slab (collective; reference HermitianSymmetrizeXY(split3&))."""
        lib.fftwpp_mpiconv3_symmetrize(self._h, ctypes.c_void_p(_ptr(array)))
        return array

    def close(self):
        if self._h:
            lib.fftwpp_mpiconv3_destroy(self._h)
            self._h = None
        if self._comm:
            lib.fftwpp_gpu_comm_destroy(self._comm)
            self._comm = ctypes.c_void_p()
        if self._comm2:
            lib.fftwpp_gpu_comm_destroy(self._comm2)
            self._comm2 = ctypes.c_void_p()


class DistributedFFT:
    """Distributed 2-D / 3-D FFT, slab decomposition (reference fft2dMPI,
    fft3dMPI, rcfft2dMPI, rcfft3dMPI; mpi/mpifftw++.h:37-585).

    N: global extents (2 or 3).  real=False: complex transform of the x x Y
    [x Z] array held as this rank's x rows; forward() leaves the X x y [x Z]
    array (this rank's y rows), sign -1 by default, unnormalised.  real=True:
    real input x x Y [x Z]; the complex output halves the last dimension and,
    in 2-D, splits that halved dimension over the ranks.  Arrays are CUDA
    tensors; complex ones need `words` complex elements of storage."""

    def __init__(self, N, rank, world, real=False, sign=-1, comm="nccl"):
        self.N, self.real, self.rank, self.world = list(N), real, rank, world
        # comm: "nccl" (bootstrap over torch.distributed) or a communicator
        # handle of fftwpp_gpu_comm_create owned by the caller
        self._own = comm == "nccl"
        self._comm = _nccl_comm(rank, world) if self._own else comm
        dims = len(self.N)
        arr = ctypes.c_size_t * dims
        self._h = lib.fftwpp_mpifft_create(1 if real else 0, dims, arr(*self.N), sign, rank,
                                           world, self._comm)
        buf = (ctypes.c_size_t * 9)()
        lib.fftwpp_mpifft_split(self._h, buf)
        self.split = dict(zip("X Y Z x y z x0 y0 z0".split(), [int(v) for v in buf]))
        self.words = int(lib.fftwpp_mpifft_words(self._h))

    def exchange_table(self, direction):
        """send/receive byte counts and displacements per peer (direction 1:
        the forward transform's exchange, 0: its inverse)"""
        n = self.world
        t = [(ctypes.c_ulonglong * n)() for _ in range(4)]
        lib.fftwpp_mpifft_exchange_table(self._h, direction, *t)
        return [[int(v) for v in a] for a in t]

    def _shape(self, a, b):
        return (a, b) if len(self.N) == 2 else (a, b, self.split["Z"])

    def input_shape(self):
        """local shape of the x-split COMPLEX data (real=True: of the
        half-spectrum as Backward leaves it before c2r)"""
        return self._shape(self.split["x"], self.split["Y"])

    def output_shape(self):
        return self._shape(self.split["X"], self.split["y"])

    def real_shape(self):
        x = self.split["x"]
        return (x, self.N[1]) if len(self.N) == 2 else (x, self.N[1], self.N[2])

    def buffer(self):
        import torch
        return torch.zeros(self.words, dtype=torch.complex128, device="cuda")

    def forward(self, src, dst=None):
        lib.fftwpp_mpifft_forward(self._h, _ptr(src), None if dst is None else _ptr(dst))
        return src if dst is None else dst

    def backward(self, src, dst=None):
        lib.fftwpp_mpifft_backward(self._h, _ptr(src), None if dst is None else _ptr(dst))
        return src if dst is None else dst

    def normalize(self, f):
        lib.fftwpp_mpifft_normalize(self._h, _ptr(f))
        return f

    def shift(self, f):
        """real=True: f *= (-1)^x (3-D: (-1)^(x+y)) on the real data, which
        centres the Fourier origin (reference Shift; Forward0 = shift + forward)"""
        lib.fftwpp_mpifft_shift(self._h, _ptr(f))
        return f

    def denyquist(self, F):
        """real=True: zero the Nyquist modes of the transformed data"""
        lib.fftwpp_mpifft_denyquist(self._h, _ptr(F))
        return F

    def close(self):
        if self._h:
            lib.fftwpp_mpifft_destroy(self._h)
            self._h = None
        if self._own and self._comm:
            lib.fftwpp_gpu_comm_destroy(self._comm)
            self._comm = ctypes.c_void_p()
