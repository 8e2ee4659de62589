"""CPU, world_size 2 and 3 over gloo: the byte counts / displacements the C++
Convolution2MPI / Convolution3MPI compute for its two exchanges (cpp/mpiconvolve.cc) must realise
the global (X x y) <-> (x x Y) block transpose of the reference's
mpitranspose localize1/localize0 (mpi/mpitranspose.h:632-931), including uneven
splits.  The exchange is emulated with torch.distributed all_to_all-style
send/recv on CPU tensors; the plan objects are the product's own (GPU plans are
lazy, so no device is needed)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from fftwpp_b200 import dist_conv

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    cases = [(3, 2, (8, 6, 4)), (3, 2, (7, 5, 3)), (3, 2, (16, 9, 2)), (3, 0, (5, 7, 3)),
             (3, 1, (8, 6, 10)),               # centred Hermitian: Z = ceil(Lz/2) modes
             (2, 0, (8, 6)), (2, 0, (9, 7))]   # Convolution2MPI: Z = 1
    for dim, fam, L in cases:
        if dim == 3:
            Lx, Ly, Lz = L
            M = [2*l for l in L] if fam != 1 else [3*l//2 for l in L]
            c = dist_conv.SlabConvolution3(Lx, Ly, Lz, *M, rank, world, family=fam,
                                           mult=2 if fam == 1 else 1, comm=None)
            assert c.split["Z"] == ((Lz + 1)//2 if fam == 1 else Lz)
        else:
            Lx, Ly = L
            c = dist_conv.SlabConvolution2(Lx, Ly, 2*Lx, 2*Ly, rank, world, comm=None)
            assert c.split["Z"] == 1
        d = c.split
        X, Y, Z = d["X"], d["Y"], d["Z"]
        ext, st = dist_conv.local_dimension(Ly, rank, world)
        assert (d["y"], d["y0"]) == (ext, st)
        ext, st = dist_conv.local_dimension(X, rank, world)
        assert (d["x"], d["x0"]) == (ext, st)
        # global field G[X][Y][Z] of complex words with a unique value per entry
        G = (np.arange(X*Y*Z, dtype=np.float64).reshape(X, Y, Z) + 1) * (1 + 0.5j)
        slab = np.ascontiguousarray(G[:, d["y0"]:d["y0"]+d["y"], :])      # X x y x Z
        sc, sd, rc, rd = c.exchange_table(0)
        sendbuf = slab.reshape(-1).view(np.uint8)
        recvbuf = np.zeros(sum(rc), dtype=np.uint8)
        reqs = []
        for p in range(world):
            if p == rank:
                recvbuf[rd[p]:rd[p]+rc[p]] = sendbuf[sd[p]:sd[p]+sc[p]]
                continue
            if sc[p]:
                reqs.append(dist.isend(torch.from_numpy(sendbuf[sd[p]:sd[p]+sc[p]].copy()), p))
        for p in range(world):
            if p != rank and rc[p]:
                t = torch.empty(rc[p], dtype=torch.uint8)
                dist.recv(t, p)
                recvbuf[rd[p]:rd[p]+rc[p]] = t.numpy()
        for r in reqs:
            r.wait()
        # unpack exactly as Convolution3MPI::transposeForward does
        T = np.zeros((d["x"], Y, Z), dtype=np.complex128)
        for p in range(world):
            py, py0 = dist_conv.local_dimension(Y, p, world)
            if py == 0 or d["x"] == 0:
                continue
            blk = recvbuf[rd[p]:rd[p]+rc[p]].view(np.complex128).reshape(d["x"], py, Z)
            T[:, py0:py0+py, :] = blk
        assert np.array_equal(T, G[d["x0"]:d["x0"]+d["x"]]), "forward transpose mismatch"
        # inverse exchange: pack, exchange, land contiguously
        sc, sd, rc, rd = c.exchange_table(1)
        sendbuf = np.zeros(sum(sc), dtype=np.uint8)
        for p in range(world):
            py, py0 = dist_conv.local_dimension(Y, p, world)
            if py == 0 or d["x"] == 0:
                continue
            blk = np.ascontiguousarray(T[:, py0:py0+py, :]).reshape(-1).view(np.uint8)
            sendbuf[sd[p]:sd[p]+sc[p]] = blk
        back = np.zeros(X*d["y"]*Z*16, dtype=np.uint8)
        reqs = []
        for p in range(world):
            if p == rank:
                back[rd[p]:rd[p]+rc[p]] = sendbuf[sd[p]:sd[p]+sc[p]]
                continue
            if sc[p]:
                reqs.append(dist.isend(torch.from_numpy(sendbuf[sd[p]:sd[p]+sc[p]].copy()), p))
        for p in range(world):
            if p != rank and rc[p]:
                t = torch.empty(rc[p], dtype=torch.uint8)
                dist.recv(t, p)
                back[rd[p]:rd[p]+rc[p]] = t.numpy()
        for r in reqs:
            r.wait()
        assert np.array_equal(back.view(np.complex128).reshape(X, d["y"], Z), slab)
        c.close()
    dist.barrier()
    dist.destroy_process_group()
    print("RANK", rank, "OK")
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_tables_realise_global_transpose(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
        assert "OK" in o
