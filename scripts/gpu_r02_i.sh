#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
FFTWPP_CONV_TMEM=0 timeout 600 python -m pytest tests/test_gpu_tma.py -q -x > gpurun_out/pytest_i.txt 2>&1; tail -15 gpurun_out/pytest_i.txt
echo "== TMEM z kernel parity"
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_baseline_configs.py -q -x -k "m512 or cfg4 or conv3d or conv1d_complex_auto" > gpurun_out/pytest_i2.txt 2>&1; tail -5 gpurun_out/pytest_i2.txt
for cfg in "X=1" "FFTWPP_CONV_TMEM=0"; do
  echo "== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_i.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'parity',d['parity']['rel_l2'],' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_i.err
done
