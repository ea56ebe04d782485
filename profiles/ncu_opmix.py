import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ie=hdr.index('Instructions Executed'); si=hdr.index('# Samples')
ops=collections.Counter(); smp=collections.Counter()
for r in rows[2:]:
    if len(r)<=ie: continue
    try: n=int(r[ie] or 0); s=int(r[si] or 0)
    except: continue
    ins=r[1].strip()
    if ins.startswith('@'): ins=ins.split(' ',1)[1].strip()
    op=ins.split(' ')[0].split('.')[0]
    full=ins.split(' ')[0]
    if op in ('LDS','STS','LDG','STG','RED','ATOMG','LDGSTS','SHFL'): op=full if op in('LDS','STS') else op
    ops[op]+=n; smp[op]+=s
tot=sum(ops.values()); ts=sum(smp.values())
for op,n in ops.most_common(28): print(f"{op:14s} {100*n/tot:5.1f}% of inst   {100*smp[op]/ts:5.1f}% of samples")
print('total inst',tot)
