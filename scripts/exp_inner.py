"""Two-stage (p > 2) 1-D convolutions: forced-m scan, device-resident, CUDA events;
parity of every variant against the default pick on random data."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftwpp_b200 as fp  # noqa: E402

rng = np.random.default_rng(9)
for lg in (16, 18, 20, 22):
    L = 1 << lg
    f = torch.from_numpy(rng.uniform(-1, 1, L) + 1j * rng.uniform(-1, 1, L)).cuda()
    g = torch.from_numpy(rng.uniform(-1, 1, L) + 1j * rng.uniform(-1, 1, L)).cuda()
    ref = None
    for m in (None, [1024], [2048], [4096], [8192]):
        if m and L // m[0] < 16:
            continue
        c = fp.HybridConv([L], [2 * L], m=m)
        d = [f.clone(), g.clone()]
        c.convolve(d)
        torch.cuda.synchronize()
        out = d[0].cpu().numpy()
        if ref is None:
            ref = out
        err = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
        n = 50 if lg <= 20 else 20
        for _ in range(3):
            c.convolve(d, normalized=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            c.convolve(d, normalized=False)
        e1.record()
        torch.cuda.synchronize()
        p = c.params(0)
        print(json.dumps({"L": "2^%d" % lg, "forced_m": m, "m": p["m"], "p": p["p"],
                          "ms": e0.elapsed_time(e1) / n, "rel_l2_vs_default": err}))
        c.close()
