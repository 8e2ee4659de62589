#!/bin/bash
# 2-GPU call: distributed parity under pytest + N=2 bench (TMA peer stores vs per-thread peer stores vs NCCL path)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_distributed.py -q > gpurun_out/pytest_g.txt 2>&1; tail -5 gpurun_out/pytest_g.txt
cat gpurun_out/dist_check_n2.log | tail -25
for cfg in "X=1" "FFTWPP_NO_TMA=1" "FFTWPP_MPI_FUSED=0"; do
  echo "== bench N=2 $cfg"
  env $cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_g.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),'conv/s',round(d['value'],1),'parity',d['parity']['rel_l2'],d['parity']['ok'],'e2e',round(d['e2e']['value'],1),' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))" || tail -5 gpurun_out/bench_g.err
done
