cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_conv.py tests/test_gpu_golden.py tests/test_gpu_pad.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pingpong',d['value'],'conv/s',d['ms_per_step'],'ms', {(k['pass'],k['op']):round(k['ms_per_step'],2) for k in d['kernels']})"
