#!/bin/bash
# long rows (m=8192) on the tensor-memory kernel: parity, timing, one ncu capture
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_long_rows.py tests/test_gpu_baseline_configs.py -x -q -s -m gpu > gpurun_out/long_p.txt 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/long_p.txt | cut -c1-400
timeout 300 python scripts/exp_long.py > gpurun_out/exp_long.jsonl 2> gpurun_out/exp_long.err; echo "exp rc=$?"
cat gpurun_out/exp_long.jsonl; tail -5 gpurun_out/exp_long.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fast_conv_rows_long -c 1 -o gpurun_out/ncu_long python scripts/exp_long.py > gpurun_out/ncu_long.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
