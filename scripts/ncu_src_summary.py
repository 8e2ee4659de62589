"""Summarise an `ncu --page source --csv` export (SASS view): dynamic
instruction mix by opcode and warp-stall samples by opcode and reason."""
import csv, sys, collections, re
csv.field_size_limit(10**9)
fn=sys.argv[1]
rows=list(csv.reader(open(fn)))
# several kernels may be concatenated: split on 'Kernel Name' rows
blocks=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name':
        cur={'name':r[1],'hdr':None,'rows':[]}; blocks.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r; continue
    cur['rows'].append(r)
for b in blocks:
    h=b['hdr']; ix={k:i for i,k in enumerate(h)}
    print('==',b['name'][:110])
    ops=collections.Counter(); samp=collections.Counter(); reasons=collections.Counter()
    stallcols=[k for k in h if k.startswith('stall_')]
    opstall=collections.defaultdict(collections.Counter)
    tot=0
    for r in b['rows']:
        src=r[ix['Source']].strip()
        m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src)
        op=m.group(2) if m else src[:12]
        op0=op.split('.')[0]
        key=op0 if op0 not in('LDS','STS','LDG','STG','BAR') else op
        n=int(r[ix['Instructions Executed']] or 0)
        s=int(r[ix['# Samples']] or 0)
        ops[key]+=n; samp[key]+=s; tot+=n
        for c in stallcols:
            v=int(r[ix[c]] or 0)
            reasons[c]+=v; opstall[key][c]+=v
    ts=sum(samp.values())
    print('total warp-instr',tot,'samples',ts)
    for k,n in ops.most_common(22):
        top=', '.join('%s %d'%(c[6:],v) for c,v in opstall[k].most_common(3) if v)
        print('  %-22s %12d (%4.1f%%)  samples %6d (%4.1f%%)  %s'%(k,n,100.0*n/tot,samp[k],100.0*samp[k]/max(ts,1),top))
    print('  stall reasons:',', '.join('%s %.1f%%'%(c[6:],100.0*v/max(sum(reasons.values()),1)) for c,v in reasons.most_common(10)))
