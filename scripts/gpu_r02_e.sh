#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tma.py tests/test_gpu_refprogs.py tests/test_gpu_wrappers.py -q > gpurun_out/pytest_e.txt 2>&1; tail -12 gpurun_out/pytest_e.txt
for cfg in "X=1" "FFTWPP_NO_TMA_REAL=1"; do
  echo "== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'parity',d['parity']['rel_l2'],' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_e.err
done
