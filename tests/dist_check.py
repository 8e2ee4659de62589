"""Run under torchrun on >= 2 GPUs: the distributed (slab, NCCL all-to-all)
convolution must equal the single-GPU convolution of the gathered input, as the
reference's mpi/tests/hybridconvr3.cc:132-167 checks (max-norm, 1e-12)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftwpp_b200 as fp  # noqa: E402
from fftwpp_b200 import dist_conv  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    fp.lib.fftwpp_gpu_set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True

    def report(tag, c, got, want_local, wmax):
        err = (np.max(np.abs(got - want_local)) if got.size else 0.0) / max(1.0, wmax)
        t = torch.tensor([err], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(tag, "ranks", world, "split", c.split, "max err", t.item())
        return t.item() < 1e-12

    def seeded(shape, seed, cplx):
        rng = np.random.default_rng(seed)       # same global field on every rank
        a = rng.uniform(-1, 1, shape)
        return a + 1j * rng.uniform(-1, 1, shape) if cplx else a

    cases = [(2, (16, 12, 20)), (2, (33, 9, 8)), (0, (8, 10, 6)), (2, (64, 64, 64))]
    if world > 4:  # every rank needs a non-empty y slab
        # uneven but non-empty slabs: Ly = 3*(world-1)+1 gives ceil-split 3,...,3,1
        cases = [(2, (16, 8 * world, 12)), (0, (9, 3 * (world - 1) + 1, 6)), (2, (64, 64, 64))]
    # the (64,64,64) case runs the fused (peer-store) exchange; repeat it on the
    # NCCL all-to-all path
    # ceil-split with an EMPTY last rank (reference localdimension,
    # mpi/mpitranspose.h:118-130): Ly = world-1 rows over world ranks
    cases += [(0, (8, world - 1, 6)), (2, (16, world - 1, 12))]
    runs = [(fam, L, None) for fam, L in cases] + [(2, (64, 64, 64), "0")]
    for fam, L, fusedenv in runs:
        M = [2 * l for l in L]
        if fusedenv is not None:
            os.environ["FFTWPP_MPI_FUSED"] = fusedenv
        c = dist_conv.SlabConvolution3(*L, *M, rank, world, family=fam)
        os.environ.pop("FFTWPP_MPI_FUSED", None)
        full = [seeded(L, 7 + a, fam == 0) for a in range(2)]
        y, y0 = c.split["y"], c.split["y0"]
        f = [torch.from_numpy(np.ascontiguousarray(a[:, y0:y0 + y, :])).cuda() for a in full]
        want = O.conv_real(full[0], full[1]) if fam == 2 else O.conv_complex(full[0], full[1])
        c.convolve(f)
        torch.cuda.synchronize()
        tag = "3-D family %d L %s%s" % (fam, L, " (NCCL path)" if fusedenv == "0" else "")
        got = f[0].cpu().numpy() if y > 0 else want[:, y0:y0 + y, :]
        ok = report(tag, c, got, want[:, y0:y0 + y, :], np.max(np.abs(want))) and ok
        c.close()

    # centred Hermitian 3-D (reference mpi/tests/hybridconvh3.cc): the global
    # field is symmetrised, every rank takes its y slice of the half-spectrum
    for L in ((8, 2 * world + 2, 10), (12, 4 * world, 7), (9, 2 * world + 1, 6)):
        M = [3 * l // 2 + 1 for l in L]
        c = dist_conv.SlabConvolution3(*L, *M, rank, world, family=fp.FAMILY_HERMITIAN,
                                       mult=fp.MULT_REALBINARY)
        H = (L[2] + 1) // 2
        full = [seeded((L[0], L[1], H), 17 + a, True) for a in range(2)]
        y, y0 = c.split["y"], c.split["y0"]
        # the distributed symmetrisation must equal the serial rule on the global field
        f = [torch.from_numpy(np.ascontiguousarray(a[:, y0:y0 + y, :])).cuda() for a in full]
        for a in full:
            O.symmetrize(L, a)
        for a in range(2):
            c.symmetrize(f[a])
            same = np.array_equal(f[a].cpu().numpy(), full[a][:, y0:y0 + y, :])
            t = torch.tensor([0.0 if same else 1.0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print("distributed HermitianSymmetrizeXY L %s array %d:" % (L, a),
                      "exact" if t.item() == 0 else "MISMATCH")
            ok = ok and t.item() == 0
        want = O.conv_hermitian(L, full[0], full[1])
        c.convolve(f)
        torch.cuda.synchronize()
        ok = report("3-D Hermitian L %s" % (L,), c, f[0].cpu().numpy(),
                    want[:, y0:y0 + y, :], np.max(np.abs(want))) and ok
        c.close()

    # pencil decomposition (forced; reference mpi/mpiconvolve.h:208-216): y over
    # py ranks, z over pz ranks, every admissible grid of this world size
    grids = [(py, world // py) for py in range(1, world + 1) if world % py == 0 and world // py > 1]
    for fam, L in ((2, (16, 12, 20)), (0, (8, 10, 6)), (2, (33, 5, 9)), (2, (64, 64, 64))):
        for (py, pz) in grids:
            M = [2 * l for l in L]
            c = dist_conv.SlabConvolution3(*L, *M, rank, world, family=fam, grid=(py, pz))
            full = [seeded(L, 57 + a, fam == 0) for a in range(2)]
            y, y0 = c.split["y"], c.split["y0"]
            z, z0 = c.zsplit["z"], c.zsplit["z0"]
            f = [torch.from_numpy(np.ascontiguousarray(a[:, y0:y0 + y, z0:z0 + z])).cuda()
                 for a in full]
            want = O.conv_real(full[0], full[1]) if fam == 2 else O.conv_complex(full[0], full[1])
            c.convolve(f)
            torch.cuda.synchronize()
            wl = want[:, y0:y0 + y, z0:z0 + z]
            got = f[0].cpu().numpy() if wl.size else wl
            ok = report("pencil %dx%d family %d L %s" % (py, pz, fam, L), c, got, wl,
                        np.max(np.abs(want))) and ok
            c.close()

    # 2-D centred Hermitian (reference mpi/tests/hybridconvh2.cc): the stored
    # modes of y (ceil(Ly/2) of them) are split over the ranks
    for L in ((8, 4 * world), (9, 4 * world + 1), (16, 6 * world - 2)):
        M = [3 * l // 2 + 1 for l in L]
        c = dist_conv.SlabConvolution2(*L, *M, rank, world, family=fp.FAMILY_HERMITIAN,
                                       mult=fp.MULT_REALBINARY)
        H = (L[1] + 1) // 2
        full = [seeded((L[0], H), 37 + a, True) for a in range(2)]
        for a in full:
            O.symmetrize(L, a)
        y, y0 = c.split["y"], c.split["y0"]
        f = [torch.from_numpy(np.ascontiguousarray(a[:, y0:y0 + y])).cuda() for a in full]
        want = O.conv_hermitian(L, full[0], full[1])
        c.convolve(f)
        torch.cuda.synchronize()
        wl = want[:, y0:y0 + y]
        got = f[0].cpu().numpy() if wl.size else wl
        ok = report("2-D Hermitian L %s" % (L,), c, got, wl, np.max(np.abs(want))) and ok
        c.close()

    # 2-D complex (reference Convolution2MPI, mpi/tests/hybridconv2.cc)
    for L in ((16, 4 * world), (33, 3 * (world - 1) + 1), (128, 64 * world)):
        M = [2 * l for l in L]
        c = dist_conv.SlabConvolution2(*L, *M, rank, world)
        full = [seeded(L, 27 + a, True) for a in range(2)]
        y, y0 = c.split["y"], c.split["y0"]
        f = [torch.from_numpy(np.ascontiguousarray(a[:, y0:y0 + y])).cuda() for a in full]
        want = O.conv_complex(full[0], full[1])
        c.convolve(f)
        torch.cuda.synchronize()
        ok = report("2-D complex L %s" % (L,), c, f[0].cpu().numpy(), want[:, y0:y0 + y],
                    np.max(np.abs(want))) and ok
        c.close()

    # distributed FFTs (reference mpi/mpifftw++.h:37-585; tests mpi/tests/fft2.cc,
    # fft3.cc, fft2r.cc, fft3r.cc): x-split input, y-split output, against numpy
    for N in ((16, 12), (4 * world, 4 * world), (9, world - 1), (world + 1, 10), (64, 64),
              (8, 6, 10), (2 * world, 3 * world, 5), (5, world + 1, 7), (64, 64, 64)):
        for real in (False, True):
            if real and N[-1] < 2:
                continue
            fft = dist_conv.DistributedFFT(N, rank, world, real=real)
            sp = fft.split
            full = seeded(N, 77, not real)
            x, x0, y, y0 = sp["x"], sp["x0"], sp["y"], sp["y0"]
            want = np.fft.rfftn(full) if real else np.fft.fftn(full)
            loc = torch.from_numpy(np.ascontiguousarray(full[x0:x0 + x])).cuda()
            if real:
                F = fft.buffer()
                fft.forward(loc, F)
            else:
                F = fft.buffer()
                F[:loc.numel()] = loc.reshape(-1)
                fft.forward(F)
            torch.cuda.synchronize()
            wl = want[:, y0:y0 + y]
            got = F.cpu().numpy()[:wl.size].reshape(wl.shape)
            tag = "%s FFT N %s" % ("real" if real else "complex", N)
            ok = report(tag + " forward", fft, got, wl, np.max(np.abs(want))) and ok
            if real:
                back = torch.zeros_like(loc)
                fft.backward(F, back)
                fft.normalize(back)
                got = back.cpu().numpy()
            else:
                fft.backward(F)
                fft.normalize(F)
                got = F.cpu().numpy()[:loc.numel()].reshape(loc.shape)
            torch.cuda.synchronize()
            ok = report(tag + " round trip", fft, got, full[x0:x0 + x], 1.0) and ok
            fft.close()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)
    if rank == 0:
        print("DIST OK")


if __name__ == "__main__":
    main()
