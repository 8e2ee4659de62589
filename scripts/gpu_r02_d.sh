#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r02d.txt 2>&1
tail -15 gpurun_out/pytest_gpu_r02d.txt
