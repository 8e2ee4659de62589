"""cfg5 (4096 rows of L=8192) on the tensor-memory long-row kernel vs the
two-stage path (FFTWPP_NO_LONG_ROWS=1), device-resident, CUDA events."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftwpp_b200 as fp  # noqa: E402

PEAK = 6536.4


def timeit(fn, n=20, warm=3):
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


def main():
    L, rows = 8192, 4096
    rng = np.random.default_rng(5)
    f = torch.from_numpy(rng.uniform(-1, 1, (rows, L)) + 1j * rng.uniform(-1, 1, (rows, L))).cuda()
    g = torch.from_numpy(rng.uniform(-1, 1, (rows, L)) + 1j * rng.uniform(-1, 1, (rows, L))).cuda()
    res = {}
    paths = (("long", None),) if os.environ.get("EXP_LONG_ONLY") else (("long", None), ("two-stage", "1"))
    for tag, env in paths:
        if env:
            os.environ["FFTWPP_NO_LONG_ROWS"] = env
        c = fp.HybridConv([L], [2 * L])
        os.environ.pop("FFTWPP_NO_LONG_ROWS", None)
        d = [f.clone(), g.clone()]
        c.convolve_rows(d, rows, L)
        torch.cuda.synchronize()
        res[tag] = d[0].clone()
        n0 = fp.lib.fftwpp_gpu_launch_count()
        c.convolve_rows(d, rows, L, normalized=False)
        launches = fp.lib.fftwpp_gpu_launch_count() - n0
        ms = timeit(lambda: c.convolve_rows(d, rows, L, normalized=False))
        gb = 3 * rows * L * 16 / 1e9
        print(json.dumps({"config": "cfg5 4096 x L=8192", "path": tag, "params": c.params(0),
                          "launches": launches, "ms": ms, "GBps": gb / (ms / 1e3),
                          "frac_of_measured_hbm": gb / (ms / 1e3) / PEAK}))
        for nb in (1, 148, 296, 1024):
            ms = timeit(lambda: c.convolve_rows(d, nb, L, normalized=False))
            print(json.dumps({"path": tag, "rows": nb, "ms": ms}))
        c.close()
    if "two-stage" not in res:
        return
    a, b = res["long"].cpu().numpy(), res["two-stage"].cpu().numpy()
    print(json.dumps({"long vs two-stage rel_l2": float(np.linalg.norm(a - b) / np.linalg.norm(b))}))


if __name__ == "__main__":
    main()
