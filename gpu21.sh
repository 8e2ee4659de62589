cd $GRAFT_REPO_ROOT
for TR in 8 16; do
FFTWPP_TILE_LANES_REAL=$TR python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2960$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e 2>&1 | grep -E '^\{|rror' | python -c "
import json,sys
t=sys.stdin.read()
try:
  d=json.loads(t); print('TR=$TR N=2',d['value'],'conv/s',d['ms_per_step'],'ms',{(k['pass'],k['op']):round(k['ms_per_step'],2) for k in d['kernels']})
except Exception as e: print('ERR',t[-1500:])"
done
FFTWPP_TILE_LANES_REAL=16 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('TR=16 N=1',d['value'],'conv/s',d['ms_per_step'],'ms', {(k['pass'],k['op']):round(k['ms_per_step'],2) for k in d['kernels']})"
