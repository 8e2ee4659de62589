cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_conv.py tests/test_gpu_golden.py tests/test_gpu_pad.py -x -q -m gpu 2>&1 | tail -2
python /dev/stdin <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
import fftwpp_b200 as fp
from oracle import oracle as O
bad=0
for L in (16,64,512,2048):
    for C,S in ((4,4),(6,6),(34,34),(8,10),(5,5)):
        for M in (2*L,3*L,4*L):
            pad=fp.Pad(3,L,M,C,S,L,1,0)
            rng=np.random.default_rng(3)
            f=np.zeros((L,S)); f[:,:C]=rng.uniform(-1,1,(L,C))
            F2=O.padded_dft(3,L,pad.paddedSize,f[:,:C])
            h=np.zeros_like(f); err=0;nrm=0
            for r in pad.residue_calls():
                F=pad.forward(np.ascontiguousarray(f),r)
                G=np.zeros_like(F)
                for k in range(pad.noutputs(r)):
                    i=pad.index(r,k)
                    val=np.array([O.real_spectrum_at(F2[:,c],pad.paddedSize,i) for c in range(C)])
                    err+=np.sum(abs(F[S*k:S*k+C]-val)**2); nrm+=np.sum(abs(val)**2); G[S*k:S*k+C]=val
                pad.backward(G,h,r)
            e1=np.sqrt(err/nrm); e2=O.rel_l2(h[:,:C]/pad.normalization,f[:,:C])
            if e1>1e-13 or e2>1e-13: print('BAD pad',L,C,S,M,e1,e2); bad+=1
for shape in ((32,32,32),(64,128,32),(128,128,128),(256,16,64),(64,6,5)):
    rng=np.random.default_rng(1); f=rng.uniform(-1,1,shape); g=rng.uniform(-1,1,shape)
    c=fp.HybridConv(list(shape),[2*s for s in shape],family=fp.FAMILY_REAL); a=[f.copy(),g.copy()]; c.convolve(a); e=O.rel_l2(a[0],O.conv_real(f,g)); print('3dr',shape,e); bad+=e>1e-13
print('BAD',bad)
PY
for st in 1 0; do for T in 4 8; do
FFTWPP_TILE_LANES=$T FFTWPP_STAGE_REAL=$st python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('stage=$st T=$T',d['value'],'conv/s',d['ms_per_step'],'ms', {(k['pass'],k['op']):round(k['ms_per_step'],2) for k in d['kernels']})
"
done; done
