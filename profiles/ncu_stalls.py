import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
sc=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
c=collections.Counter()
for r in rows[2:]:
    for i,h in sc:
        try: c[h]+=int(r[i] or 0)
        except: pass
t=sum(c.values())
for h,n in c.most_common(): print(f"{h:24s} {100*n/t:5.1f}%")
