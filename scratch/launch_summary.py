import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    name=d['Kernel Name'].split('(')[0]; v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
    v*= {'ns':1,'us':1e3,'ms':1e6,'s':1e9}.get(u,1)
    a=agg.setdefault(name,[0,0.0,0.0]); a[0]+=1; a[1]+=v; a[2]=max(a[2],v)
tot=sum(a[1] for a in agg.values())
print(f"{'kernel':34s} {'n':>4s} {'total us':>10s} {'mean us':>9s} {'max us':>9s} {'share':>6s}")
for k,(n,t,m) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:34s} {n:4d} {t/1e3:10.1f} {t/n/1e3:9.1f} {m/1e3:9.1f} {100*t/tot:5.1f}%")
