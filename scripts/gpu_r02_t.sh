#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python scripts/exp_inner.py > gpurun_out/exp_inner.jsonl 2> gpurun_out/exp_inner.err; echo "exp rc=$?"
cat gpurun_out/exp_inner.jsonl; tail -3 gpurun_out/exp_inner.err
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_baseline_configs.py tests/test_gpu_long_rows.py -x -q -m gpu > gpurun_out/pytest_t.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_t.txt | cut -c1-300
