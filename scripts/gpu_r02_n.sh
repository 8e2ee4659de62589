#!/bin/bash
# 8-GPU call: wide (128-byte row) remote tiles of the y backward pass, A/B
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for cfg in "X=1" "FFTWPP_TMA_WIDE_REMOTE=0"; do
  echo "== bench N=$N $cfg"
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_n.err | tail -1 > gpurun_out/bench_n_$N.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n_$N.json').read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'conv/s',round(d['value'],1),'parity',d['parity']['rel_l2'],d['parity']['ok'],' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_n.err
done
