#!/bin/bash
# round-2 GPU call C: TMA strided kernels -- parity and A/B timing
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== swizzled tiles"; timeout 600 python -m pytest tests/test_gpu_tma.py -q -x 2>&1 | tail -12
echo "== dense tiles"; FFTWPP_TMA_SWIZZLE=0 timeout 600 python -m pytest tests/test_gpu_tma.py -q -x 2>&1 | tail -12
for cfg in "X=1" "FFTWPP_TMA_SWIZZLE=0" "FFTWPP_NO_TMA=1"; do
  echo "== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'parity',d['parity']['rel_l2'],' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_c.err
done
echo "== full suite"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8
