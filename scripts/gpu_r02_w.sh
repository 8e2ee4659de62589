#!/bin/bash
# rows of 2048 points on the tensor-memory kernel (experimental build lib_fftwpp_lg11.so) vs fast_conv_rows_q2
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_long_rows.py tests/test_gpu_baseline_configs.py -x -q -m gpu 2>&1 | tail -2
FFTWPP_LIB=lib_fftwpp_lg11.so timeout 200 python -m pytest tests/test_gpu_long_rows.py tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -2
(python scripts/exp_rows.py 2048 16384; FFTWPP_LIB=lib_fftwpp_lg11.so python scripts/exp_rows.py 2048 16384; FFTWPP_LIB=lib_fftwpp_lg11.so python scripts/exp_rows.py 2048 1) > gpurun_out/exp_rows_w.jsonl 2>&1
cat gpurun_out/exp_rows_w.jsonl | cut -c1-400
