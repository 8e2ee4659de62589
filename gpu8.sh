cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -3
python /dev/stdin <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
import fftwpp_b200 as fp
from oracle import oracle as O
for L in (16,32,64,128,256,512,1024,2048,4096):
    rng=np.random.default_rng(L); f=rng.uniform(-1,1,L)+1j*rng.uniform(-1,1,L); g=rng.uniform(-1,1,L)+1j*rng.uniform(-1,1,L)
    for M,m in ((2*L,None),(2*L,[L//2]),(3*L,[L]),(5*L,[L])):
        c=fp.HybridConv([L],[M],m=m,D=[1] if m else None,I=[0] if m else None); a=[f.copy(),g.copy()]; c.convolve(a)
        e=O.rel_l2(a[0],O.conv_complex(f,g))
        print('1d',L,M,c.params(0)['m'],c.params(0)['q'],e, '' if e<1e-13 else 'BAD')
for shape in ((128,256),(512,512)):
    rng=np.random.default_rng(1); f=rng.uniform(-1,1,shape)+1j*rng.uniform(-1,1,shape); g=rng.uniform(-1,1,shape)+1j*rng.uniform(-1,1,shape)
    c=fp.HybridConv(list(shape),[2*s for s in shape]); a=[f.copy(),g.copy()]; c.convolve(a); print('2d',shape,O.rel_l2(a[0],O.conv_complex(f,g)))
for shape in ((64,128,32),(128,128,128)):
    rng=np.random.default_rng(1); f=rng.uniform(-1,1,shape); g=rng.uniform(-1,1,shape)
    c=fp.HybridConv(list(shape),[2*s for s in shape],family=fp.FAMILY_REAL); a=[f.copy(),g.copy()]; c.convolve(a); print('3dr',shape,O.rel_l2(a[0],O.conv_real(f,g)))
PY
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],'conv/s',d['ms_per_step'],'ms', 'conv frac',d['roofline_conv']['frac'])
for k in d['kernels']: print('  ',k['pass'],k['op'],round(k['ms_per_step'],3),'ms',k['launches_per_step'],round(k['GBps'] or 0,1),'GB/s')
"
