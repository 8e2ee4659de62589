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
    cases = [(2, (16, 12, 20)), (2, (33, 9, 8)), (0, (8, 10, 6)), (2, (64, 64, 64))]
    if world > 4:  # every rank needs a non-empty y slab
        # uneven but non-empty slabs: Ly = 3*(world-1)+1 gives ceil-split 3,...,3,1
        cases = [(2, (16, 8 * world, 12)), (0, (9, 3 * (world - 1) + 1, 6)), (2, (64, 64, 64))]
    for fam, L in cases:
        M = [2 * l for l in L]
        c = dist_conv.SlabConvolution3(*L, *M, rank, world, family=fam)
        f = c.make_inputs(seed=7, scale_second=1.0)
        # gather the global inputs on every rank (test only)
        full = []
        for a in range(2):
            parts = [None] * world
            dist.all_gather_object(parts, f[a].cpu().numpy())
            full.append(np.concatenate(parts, axis=1))
        want = O.conv_real(full[0], full[1]) if fam == 2 else O.conv_complex(full[0], full[1])
        c.convolve(f)
        torch.cuda.synchronize()
        y, y0 = c.split["y"], c.split["y0"]
        got = f[0].cpu().numpy()
        ref = want[:, y0:y0 + y, :]
        err = np.max(np.abs(got - ref)) / max(1.0, np.max(np.abs(want)))
        t = torch.tensor([err], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("family", fam, "L", L, "ranks", world, "split", c.split, "max err", t.item())
        ok = ok and t.item() < 1e-12
        c.close()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)
    if rank == 0:
        print("DIST OK")


if __name__ == "__main__":
    main()
