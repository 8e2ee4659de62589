#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tests/dist_check.py > gpurun_out/dist_l_n$N.log 2>&1; echo "rc=$?"
grep -v "^\[W\|^$\|\*\*\*\|OMP_NUM\|^W1017" gpurun_out/dist_l_n$N.log | cut -c1-230 | tail -45
