import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
agg=collections.defaultdict(lambda:[0,0,collections.Counter(),''])
fname=''
hdr=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': fname=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; si=hdr.index('# Samples'); ie=hdr.index('Instructions Executed'); sc=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]; continue
    if hdr is None or len(r)<=si: continue
    try: n=int(r[si] or 0)
    except: continue
    key=(fname,r[0])
    a=agg[key]; a[0]+=n
    try: a[1]+=int(r[ie] or 0)
    except: pass
    for i,h in sc:
        try: a[2][h]+=int(r[i] or 0)
        except: pass
    if r[1]: a[3]=r[1][:110]
tot=sum(a[0] for a in agg.values())
print('total samples',tot)
top=sorted(agg.items(),key=lambda kv:-kv[1][0])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]
for (f,l),a in top:
    print(f"{100*a[0]/tot:5.1f}% inst={a[1]:>10} {f}:{l:>4} {[(h.replace('stall_',''),c) for h,c in a[2].most_common(3)]}  | {a[3].strip()}")
