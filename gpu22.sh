cd $GRAFT_REPO_ROOT
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('prefetch',d['value'],'conv/s',d['ms_per_step'],'ms', {(k['pass'],k['op']):round(k['ms_per_step'],2) for k in d['kernels']})"
