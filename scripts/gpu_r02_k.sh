#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for mode in 2 1 0; do
  echo "== FFTWPP_CONV_TMEM=$mode"
  FFTWPP_CONV_TMEM=$mode timeout 300 python -m pytest -s -q tests/test_gpu_conv.py -k "m512_batches or conv3d" 2>&1 | tail -2
  FFTWPP_CONV_TMEM=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_k.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'parity',d['parity']['rel_l2'],' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_k.err
done
