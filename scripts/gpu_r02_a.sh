#!/bin/bash
# round-2 GPU call A: new parity tests + bench line with parity + tile-width A/B
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
python -c "import os; print('cores', len(os.sched_getaffinity(0)))"
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_r02a.txt
cat gpurun_out/pytest_gpu_r02a.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02a_n1.json 2> gpurun_out/bench_r02a_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02a_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'parity',d.get('parity'),'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'])
for k in d['kernels']: print(k['pass'],k['op'],round(k['ms_per_step'],3))
PY
for cfg in "FFTWPP_TILE_LANES=8" "FFTWPP_TILE_LANES=2" "FFTWPP_TILE_LANES_REAL=16" "FFTWPP_TILE_LANES_REAL=4" "FFTWPP_TILE_LANES=8 FFTWPP_TILE_LANES_REAL=16" "FFTWPP_PINGPONG=1"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms',round(d['ms_per_step'],3),' '.join('%s-%s=%.3f'%(k['pass'],k['op'][:3],k['ms_per_step']) for k in d['kernels']))"
done
