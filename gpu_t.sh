cd $GRAFT_REPO_ROOT
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 N=1',d['value'],'conv/s',d['ms_per_step'],'ms', {(k['pass'],k['op']):round(k['ms_per_step'],2) for k in d['kernels']})"; }
run default
FFTWPP_GATHER_FWD=1 run gather
FFTWPP_DIRECT3=1 run direct3
FFTWPP_GATHER_FWD=1 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "complex or 2d or 3d" 2>&1 | tail -2
FFTWPP_DIRECT3=1 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "complex or 2d or 3d" 2>&1 | tail -2
