#!/bin/bash
# full GPU suite with the long-row kernel serving m=4096 and m=8192, then A/B timings
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02r.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_r02r.txt | cut -c1-300
(
python scripts/exp_rows.py 4096 8192
FFTWPP_LONG_MIN_LG=13 python scripts/exp_rows.py 4096 8192
python scripts/exp_rows.py 8192 4096
FFTWPP_LONG_PREFETCH=0 python scripts/exp_rows.py 8192 4096
FFTWPP_LONG_PREFETCH=0 python scripts/exp_rows.py 4096 8192
python scripts/exp_cfg2.py
FFTWPP_LONG_MIN_LG=13 python scripts/exp_cfg2.py
) > gpurun_out/exp_rows_r.jsonl 2> gpurun_out/exp_rows_r.err
cat gpurun_out/exp_rows_r.jsonl | cut -c1-700; tail -3 gpurun_out/exp_rows_r.err
