#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for promo in 0 2 3 1; do
  echo "== L2 promotion $promo"
  FFTWPP_LIB=lib_fftwpp_exp.so FFTWPP_TMA_L2PROMO=$promo timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_m.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'parity',d['parity']['rel_l2'],' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_m.err
done
