#!/bin/bash
# long rows: 16x512 engine (default) vs the RegFFT<13> engine; parity, timing, ncu
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_long_rows.py tests/test_gpu_baseline_configs.py -x -q -s -m gpu > gpurun_out/long_q.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/long_q.txt | cut -c1-300
FFTWPP_LONG_ENGINE=1 timeout 600 python -m pytest tests/test_gpu_long_rows.py -x -q -s -m gpu > gpurun_out/long_q1.txt 2>&1; echo "pytest(engine 1) rc=$?"
tail -4 gpurun_out/long_q1.txt | cut -c1-300
timeout 300 python scripts/exp_long.py > gpurun_out/exp_long2.jsonl 2> gpurun_out/exp_long2.err; echo "exp rc=$?"
cat gpurun_out/exp_long2.jsonl; tail -5 gpurun_out/exp_long2.err
EXP_LONG_ONLY=1 FFTWPP_LONG_ENGINE=1 timeout 300 python scripts/exp_long.py > gpurun_out/exp_long1.jsonl 2> gpurun_out/exp_long1.err; echo "exp(engine 1) rc=$?"
cat gpurun_out/exp_long1.jsonl
EXP_LONG_ONLY=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fast_conv_rows_long -c 1 -o gpurun_out/ncu_long2 python scripts/exp_long.py > gpurun_out/ncu_long2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
