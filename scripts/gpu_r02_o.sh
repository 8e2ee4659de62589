#!/bin/bash
# f4 validation: single-rank FFT classes under pytest, then the torchrun check
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_mpifft.py -x -q -s -m gpu > gpurun_out/mpifft_o.txt 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/mpifft_o.txt | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tests/dist_check.py > gpurun_out/dist_o_n$N.log 2>&1; echo "rc=$?"
grep -v "^\[W\|^$\|\*\*\*\|OMP_NUM\|^W1017" gpurun_out/dist_o_n$N.log | cut -c1-200 | tail -48
