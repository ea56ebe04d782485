mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small_batch_launches.csv python profiles/small_batch_kernels.py > gpurun_out/r02_small_batch_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_small_batch_launches.csv')))
hdr=None; tot=0; n=0
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
            us = v/1000 if u in('ns','nsecond') else v if u in ('us','usecond') else v*1000
            print(f"{us:8.1f} us  {d['Kernel Name'][:90]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
            tot+=us; n+=1
print('kernels',n,'sum us',round(tot,1))
PY
