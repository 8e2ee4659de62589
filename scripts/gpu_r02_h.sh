#!/bin/bash
# 8-GPU call: distributed parity at N=8 under pytest + bench at N=8 and N=4
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_distributed.py -q > gpurun_out/pytest_h.txt 2>&1; tail -4 gpurun_out/pytest_h.txt
tail -3 gpurun_out/dist_check_n8.log | head -2
for n in 8 4; do
  echo "== bench N=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/bench_h$n.err | tail -1 > gpurun_out/bench_r02_n$n.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_r02_n$n.json').read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'conv/s',round(d['value'],1),'parity',d['parity']['rel_l2'],d['parity']['ok'],'e2e',round(d['e2e']['value'],1),'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'],d['cpu_baseline'] and d['cpu_baseline']['cores'],' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_h$n.err
done
