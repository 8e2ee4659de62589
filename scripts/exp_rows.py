"""Time `rows` fused 1-D convolutions of length L (M=2L) in one batched launch."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftwpp_b200 as fp  # noqa: E402

L, rows = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(5)
d = [torch.from_numpy(rng.uniform(-1, 1, (rows, L)) + 1j * rng.uniform(-1, 1, (rows, L))).cuda()
     for _ in range(2)]
c = fp.HybridConv([L], [2 * L])
for _ in range(3):
    c.convolve_rows(d, rows, L, normalized=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    c.convolve_rows(d, rows, L, normalized=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
gb = 3 * rows * L * 16 / 1e9
print(json.dumps({"L": L, "rows": rows, "params": c.params(0), "ms": ms, "GBps": gb / (ms / 1e3),
                  "env": {k: v for k, v in os.environ.items() if k.startswith("FFTWPP_")}}))
