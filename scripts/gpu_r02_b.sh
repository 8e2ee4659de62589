#!/bin/bash
# round-2 GPU call B: re-run new tests, calibration microbench, ncu source-level captures
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_refprogs.py tests/test_gpu_custom_mult.py tests/test_gpu_wrappers.py -q 2>&1 | tail -8
./scripts/ubench/ubench > gpurun_out/ubench_r02.txt 2>&1; cat gpurun_out/ubench_r02.txt
for spec in fast_conv_rows_pipe:1:1 fast_forward_many:5:2 fast_backward_many:2:2; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -f -o /tmp/src_$k \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
  ls -la /tmp/src_$k.ncu-rep
  ncu -i /tmp/src_$k.ncu-rep --page source --csv > gpurun_out/ncu_src_$k.csv 2>/dev/null
  ncu -i /tmp/src_$k.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$k.csv 2>/dev/null
done
ls -la gpurun_out | tail
