"""Timing of BASELINE cfg2 (2-D complex 4096^2) with the chooser's own parameters,
per-pass event profile included; prints one JSON line like profiles/run_configs.py."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftwpp_b200 as fp  # noqa: E402
from oracle import oracle as O  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n = 4096
rng = np.random.default_rng(1234)
c = fp.HybridConv([n, n], [2 * n, 2 * n])
d = [torch.zeros((n, n), dtype=torch.complex128, device="cuda") for _ in range(2)]
ms = timeit(lambda: c.convolve(d, normalized=False))
fp.profile_enable(True)
c.convolve(d, normalized=False)
torch.cuda.synchronize()
prof = {"%s %s" % k: v[0] for k, v in fp.profile_read().items()}
fp.profile_enable(False)
f = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
g = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
a = [f.copy(), g.copy()]
c.convolve(a)
err = O.rel_l2(a[0], O.conv_complex(f, g))
gb = 15 * n * n * 16 / 1e9
print(json.dumps({"config": "cfg2 2-D complex 4096^2", "ms": ms, "conv_per_s": 1e3 / ms,
                  "sweep_model_GB": gb, "GBps": gb / (ms / 1e3),
                  "frac_of_measured_hbm": gb / (ms / 1e3) / 6536.4, "rel_l2_vs_oracle": err,
                  "params": [c.params(0), c.params(1)], "pass_ms": prof}))
