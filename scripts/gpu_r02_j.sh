#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
FFTWPP_CONV_TMEM=0 timeout 600 python -u -X faulthandler -m pytest -s -q tests/test_gpu_tma.py > gpurun_out/pytest_j.txt 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_j.txt
