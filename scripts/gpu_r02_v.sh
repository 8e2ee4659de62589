#!/bin/bash
# N=2 bench line on the final build (wide remote tiles in the y backward pass)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; echo "bench rc=$?"
grep '^{' gpurun_out/bench_r02_n$N.json | cut -c1-2500
tail -3 gpurun_out/bench_r02_n$N.err | cut -c1-300
