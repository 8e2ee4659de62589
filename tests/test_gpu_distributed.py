"""Distributed (slab over y) convolutions under pytest: when the box has two
or more GPUs, launch tests/dist_check.py with torchrun and require its
`DIST OK` (reference protocol mpi/tests/hybridconvr3.cc:132-167: distributed
result == serial convolution of the gathered input, max-norm 1e-12)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import fftwpp_b200 as fp
    return fp.lib.fftwpp_gpu_device_count()


def _run(nproc, port, extra_env=None):
    env = dict(os.environ)
    env.pop("NCCL_DEBUG", None)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        with open(os.path.join(log, "dist_check_n%d.log" % nproc), "w") as fh:
            fh.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    return r


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_distributed_equals_serial(nproc):
    n = _ngpu()
    if n < nproc:
        pytest.skip("needs %d GPUs, box has %d" % (nproc, n))
    if nproc not in (2, n):
        pytest.skip("covered by the 2-GPU and the full-box runs")
    r = _run(nproc, 29610 + nproc)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST OK" in r.stdout
