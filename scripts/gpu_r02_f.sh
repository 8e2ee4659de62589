#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_f.txt 2>&1; tail -6 gpurun_out/pytest_f.txt
for cfg in "X=1" "FFTWPP_CONV_Q2=0"; do
  echo "== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_f.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'parity',d['parity']['rel_l2'],' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_f.err
done
timeout 600 python profiles/run_configs.py > gpurun_out/configs_r02f.jsonl 2> gpurun_out/configs_r02f.err; cut -c1-230 gpurun_out/configs_r02f.jsonl
